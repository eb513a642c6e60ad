#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_postproc.py -q > gpurun_out/run26_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/run26_pytest.log
tail -3 gpurun_out/run26_pytest.log | cut -c1-200
timeout 120 python tools/layer_bench.py --calibrated --reps 3 --tag r26_cand --only cand 2>&1 | grep cand | head -5 | cut -c1-110
for pdl in 0 1; do
ODTK_PDL=$pdl timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e > gpurun_out/run26_bench_$pdl.json 2>> gpurun_out/run26_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/run26_bench_$pdl.json").read().strip().splitlines()[-1]); print("pdl=$pdl", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"], d["latency"])
PY
done
