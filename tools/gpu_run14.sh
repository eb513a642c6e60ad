#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -q -x -k "bottleneck" > gpurun_out/run14_bt.log 2>&1; echo "rc=$?" >> gpurun_out/run14_bt.log
tail -5 gpurun_out/run14_bt.log | cut -c1-200
ODTK_BNECK_PITCH=16 timeout 300 python -m pytest tests/test_gpu_conv.py -q -x -k "bottleneck" > gpurun_out/run14_bt16.log 2>&1; echo "rc=$?" >> gpurun_out/run14_bt16.log
tail -3 gpurun_out/run14_bt16.log | cut -c1-200
timeout 300 python tools/layer_bench.py --tag r14_base --only bneck > gpurun_out/run14_lb_base.log 2>&1
ODTK_BNECK_PITCH=16 timeout 300 python tools/layer_bench.py --tag r14_p16 --only bneck > gpurun_out/run14_lb_p16.log 2>&1
ODTK_BNECK_NR=2 timeout 300 python tools/layer_bench.py --tag r14_nr2 --only bneck > gpurun_out/run14_lb_nr2.log 2>&1
ODTK_BNECK_NR=3 timeout 300 python tools/layer_bench.py --tag r14_nr3 --only bneck > gpurun_out/run14_lb_nr3.log 2>&1
grep -h bneck gpurun_out/run14_lb_*.log | cut -c1-120
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -q > gpurun_out/run14_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/run14_pytest.log
tail -5 gpurun_out/run14_pytest.log | cut -c1-200
ODTK_FUSED_BNECK=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e > gpurun_out/run14_bench_unfused.json 2> gpurun_out/run14_bench.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e > gpurun_out/run14_bench.json 2>> gpurun_out/run14_bench.err
python - <<'PY'
import json
for f in ("run14_bench_unfused","run14_bench"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"])
    except Exception as e: print(f, "failed", e)
PY
