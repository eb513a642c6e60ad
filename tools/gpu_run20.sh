#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/run20_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/run20_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/run20_pytest.log | head
ODTK_FUSED_BNECK=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e > gpurun_out/run20_bench_unfused.json 2> gpurun_out/run20_bench.err
ODTK_BENCH_INSTEP=gpurun_out/run20_instep.json timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e > gpurun_out/run20_bench.json 2>> gpurun_out/run20_bench.err
python - <<'PY'
import json
for f in ("run20_bench_unfused","run20_bench"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"])
    except Exception as e: print(f, "failed", e)
PY
