#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_postproc.py -q 2>&1 | tail -2
for st in 1 2 3; do
ODTK_BENCH_POSTPROC_STREAMS=$st timeout 300 python bench.py --config postproc --steps 100 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/run28_pp_$st.json 2> gpurun_out/run28_pp_$st.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/run28_pp_$st.json").read().strip().splitlines()[-1]); print("streams=$st", d["value"], "img/s", d["us_per_image"], "us/img", d["roofline"]["achieved"], d["roofline"]["frac"])
except Exception as e: print("streams=$st failed", e, open("gpurun_out/run28_pp_$st.err").read()[-500:])
PY
done
ODTK_BENCH_POSTPROC_STREAMS=2 timeout 300 python bench.py --config postproc_rotated --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rotated streams=2', d['value'], d['us_per_image'])"
