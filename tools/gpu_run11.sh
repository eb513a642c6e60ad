#!/bin/bash
set -x
mkdir -p gpurun_out
ODTK_BENCH_INSTEP=gpurun_out/run11_instep.json timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e > gpurun_out/run11_bench.json 2> gpurun_out/run11_bench.err
tail -c 300 gpurun_out/run11_bench.json
