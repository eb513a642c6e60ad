#!/bin/bash
# A/B of the cheap conv tweaks (layer_bench, isolated launches) + conv tests
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -q -x > gpurun_out/run12_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/run12_pytest.log
tail -3 gpurun_out/run12_pytest.log
LB="timeout 300 python tools/layer_bench.py"
$LB --tag r12_base > gpurun_out/run12_lb_base.log 2>&1
ODTK_CONV_CAND_T=0 $LB --tag r12_candt0 --only cand > gpurun_out/run12_lb_candt0.log 2>&1
ODTK_CONV_TWO_128=0 $LB --tag r12_two128_0 --only "128->128" > gpurun_out/run12_lb_two128_0.log 2>&1
ODTK_CONV_KHEAVY=1 $LB --tag r12_kheavy1 --only "2048->256 s2" > gpurun_out/run12_lb_kheavy1.log 2>&1
ODTK_CONV_KHEAVY=2 $LB --tag r12_kheavy2 --only "2048->256 s2" > gpurun_out/run12_lb_kheavy2.log 2>&1
ODTK_CONV_DEEP_BIAS_EPI=1 $LB --tag r12_deepbias --only "conv1x1" > gpurun_out/run12_lb_deepbias.log 2>&1
ODTK_B200_LIB=tools/_ab/lib_stem8.so $LB --tag r12_stem8 --only "stem" > gpurun_out/run12_lb_stem8.log 2>&1
grep -h "stem\|2048->256 s2\|cand\|128->128" gpurun_out/run12_lb_*.log | cut -c1-170
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e > gpurun_out/run12_bench.json 2> gpurun_out/run12_bench.err
tail -c 300 gpurun_out/run12_bench.json
