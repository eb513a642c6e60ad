#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -q -x -k "bottleneck" > gpurun_out/run22_bt.log 2>&1; echo "rc=$?" >> gpurun_out/run22_bt.log
tail -3 gpurun_out/run22_bt.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -q > gpurun_out/run22_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/run22_pytest.log
tail -3 gpurun_out/run22_pytest.log | cut -c1-200
timeout 300 python tools/layer_bench.py --tag r22_base > gpurun_out/run22_lb_base.log 2>&1
head -24 gpurun_out/run22_lb_base.log | cut -c1-150; tail -1 gpurun_out/run22_lb_base.log
ODTK_BENCH_INSTEP=gpurun_out/run22_instep.json timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e > gpurun_out/run22_bench.json 2>> gpurun_out/run22_bench.err
python - <<'PY'
import json
for f in ("run22_bench",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"])
    except Exception as e: print(f, "failed", e)
PY
