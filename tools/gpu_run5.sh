#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/run5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/run5_pytest.log
tail -8 gpurun_out/run5_pytest.log
timeout 300 python tools/layer_bench.py --tag r5stem --only stem > gpurun_out/run5_lb_stem.log 2>&1
ODTK_BENCH_INSTEP=gpurun_out/run5_instep.json timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/run5_bench.json 2> gpurun_out/run5_bench.err
tail -c 1500 gpurun_out/run5_bench.json; tail -5 gpurun_out/run5_bench.err
