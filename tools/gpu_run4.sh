#!/bin/bash
# round-2 GPU run 4: shared-space epilogue accesses (LDS/STS instead of generic), gather tests, full-size parity tests
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/run4_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/run4_pytest.log
tail -15 gpurun_out/run4_pytest.log
LB="timeout 300 python tools/layer_bench.py"
$LB --tag r4base > gpurun_out/run4_lb_base.log 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/run4_bench.json 2> gpurun_out/run4_bench.err
tail -c 300 gpurun_out/run4_bench.json; tail -5 gpurun_out/run4_bench.err
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:stem_pool -c 1 -f -o gpurun_out/run4_ncu_stempool python tools/capture_step.py > gpurun_out/run4_ncu_stempool.log 2>&1
