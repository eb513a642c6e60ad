#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -q -x -k "bottleneck" > gpurun_out/run15_bt.log 2>&1; echo "rc=$?" >> gpurun_out/run15_bt.log
tail -5 gpurun_out/run15_bt.log | cut -c1-200
timeout 300 python tools/layer_bench.py --tag r15_base --only bneck > gpurun_out/run15_lb_base.log 2>&1
ODTK_BNECK_NR=2 timeout 300 python tools/layer_bench.py --tag r15_nr2 --only bneck > gpurun_out/run15_lb_nr2.log 2>&1
grep -h bneck gpurun_out/run15_lb_*.log | cut -c1-120
ODTK_FUSED_BNECK=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e > gpurun_out/run15_bench_unfused.json 2> gpurun_out/run15_bench.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e > gpurun_out/run15_bench.json 2>> gpurun_out/run15_bench.err
python - <<'PY'
import json
for f in ("run15_bench_unfused","run15_bench"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"])
    except Exception as e: print(f, "failed", e)
PY
