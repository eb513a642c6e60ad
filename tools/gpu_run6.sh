#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/run6_pytest_model.log 2>&1; echo "rc=$?" >> gpurun_out/run6_pytest_model.log
grep -E "AssertionError|passed|failed" gpurun_out/run6_pytest_model.log | head -20
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_cases.py > gpurun_out/run6_sanitizer_$tool.log 2>&1; echo "rc=$?" >> gpurun_out/run6_sanitizer_$tool.log
  tail -4 gpurun_out/run6_sanitizer_$tool.log
done
timeout 300 python bench.py --config postproc --steps 50 --no-cpu-baseline > gpurun_out/run6_postproc.json 2> gpurun_out/run6_postproc.err
ODTK_FILTER_VEC=4 timeout 300 python bench.py --config postproc --steps 50 --no-cpu-baseline > gpurun_out/run6_postproc_v4.json 2>> gpurun_out/run6_postproc.err
tail -c 700 gpurun_out/run6_postproc.json; tail -c 700 gpurun_out/run6_postproc_v4.json
