#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_postproc.py -q > gpurun_out/run23_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/run23_pytest.log
tail -3 gpurun_out/run23_pytest.log | cut -c1-200
ODTK_BENCH_INSTEP=gpurun_out/run23_instep.json timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e > gpurun_out/run23_bench.json 2>> gpurun_out/run23_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/run23_bench.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"])
d=json.load(open('gpurun_out/run23_instep.json'))
print(d['sum_us'])
for r in d['rows'][:40]: print("%-46s n=%-2d %8.1f us %5.1f%% %7.1f TF %7.1f GB/s"%(r['layer'][:46],r['n'],r['us'],100*r['share'],r['tflops'],r['gbs']))
PY
