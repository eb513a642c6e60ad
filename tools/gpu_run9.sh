#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
timeout 120 python tools/bisect_stem.py retinanet-examples_b200/libodtk_b200.so > gpurun_out/run9_stem.log 2>&1
ODTK_STEM_ROWS=0 timeout 120 python tools/bisect_stem.py retinanet-examples_b200/libodtk_b200.so >> gpurun_out/run9_stem.log 2>&1
cat gpurun_out/run9_stem.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/run9_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/run9_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/run9_pytest.log | head -40
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/run9_bench.json 2> gpurun_out/run9_bench.err
tail -c 3000 gpurun_out/run9_bench.json; tail -5 gpurun_out/run9_bench.err
