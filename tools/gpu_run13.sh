#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -q -x -k "bottleneck" > gpurun_out/run13_bt.log 2>&1; echo "rc=$?" >> gpurun_out/run13_bt.log
tail -30 gpurun_out/run13_bt.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -q > gpurun_out/run13_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/run13_pytest.log
tail -5 gpurun_out/run13_pytest.log | cut -c1-200
timeout 300 python tools/layer_bench.py --tag r13_base > gpurun_out/run13_lb_base.log 2>&1
head -30 gpurun_out/run13_lb_base.log | cut -c1-150
ODTK_FUSED_BNECK=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e > gpurun_out/run13_bench_unfused.json 2> gpurun_out/run13_bench.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e > gpurun_out/run13_bench.json 2>> gpurun_out/run13_bench.err
python - <<'PY'
import json
for f in ("run13_bench_unfused","run13_bench"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"])
    except Exception as e: print(f, "failed", e)
PY
tail -5 gpurun_out/run13_bench.err
