#!/bin/bash
# round-2 GPU run 1: regression tests, per-layer A/B of the env toggles, baseline bench, ncu captures of the weak layers
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/run1_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/run1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/run1_pytest.log
tail -5 gpurun_out/run1_pytest.log
timeout 300 python tools/layer_bench.py --tag base > gpurun_out/run1_lb_base.log 2>&1
ODTK_STEM_ROWS=0 timeout 300 python tools/layer_bench.py --tag stemrows0 --only stem > gpurun_out/run1_lb_stemrows0.log 2>&1
ODTK_CONV_CLUSTER_RES=1 timeout 300 python tools/layer_bench.py --tag clres1 --only res > gpurun_out/run1_lb_clres1.log 2>&1
ODTK_CONV_CLUSTER_RES=2 timeout 300 python tools/layer_bench.py --tag clres2 --only res > gpurun_out/run1_lb_clres2.log 2>&1
ODTK_CONV_CLUSTER_1X1=1 timeout 300 python tools/layer_bench.py --tag cl1x1 --only conv1x1 > gpurun_out/run1_lb_cl1x1.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/run1_bench.json 2> gpurun_out/run1_bench.err
tail -c 600 gpurun_out/run1_bench.json
for spec in "70 boxfinal" "27 res1024" "55 lateral3" "3 conv64" ; do
  set -- $spec
  timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_gemm -s $1 -c 1 -f -o gpurun_out/run1_ncu_$2 python tools/capture_step.py > gpurun_out/run1_ncu_$2.log 2>&1
done
ls -la gpurun_out | tail -30
