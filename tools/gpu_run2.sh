#!/bin/bash
# round-2 GPU run 2: new kernels (fused stem+pool, tensor-core upsample-add, element-strided stride 2), A/B of the toggles
set -x
mkdir -p gpurun_out
for t in test_stem_pool_fused_equals_stem_then_maxpool test_pad_input_rows_kernel test_upsample_add test_strided_conv_odd_sizes test_relu_copy test_conv_variant_selection test_tensor_map_cache; do
  timeout 180 python -m pytest tests/test_gpu_conv.py -q -k $t > gpurun_out/run2_t_$t.log 2>&1; echo "$t rc=$?" >> gpurun_out/run2_tests.txt
done
cat gpurun_out/run2_tests.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/run2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/run2_pytest.log
tail -15 gpurun_out/run2_pytest.log
LB="timeout 300 python tools/layer_bench.py"
$LB --tag r2base > gpurun_out/run2_lb_base.log 2>&1
ODTK_FUSED_STEM=0 $LB --tag r2_stem_unfused --only stem > gpurun_out/run2_lb_stem_unfused.log 2>&1
ODTK_FUSED_STEM=0 ODTK_STEM_ROWS=0 $LB --tag r2_stem_rows0 --only stem > gpurun_out/run2_lb_stem_rows0.log 2>&1
ODTK_CONV_UP_MMA=0 $LB --tag r2_upmma0 --only "+up" > gpurun_out/run2_lb_upmma0.log 2>&1
ODTK_CONV_MAX_STAGES=8 $LB --tag r2_st8 --only "256->" > gpurun_out/run2_lb_st8.log 2>&1
ODTK_CONV_TWO_NARROW=1 $LB --tag r2_two_narrow --only f32 > gpurun_out/run2_lb_two_narrow.log 2>&1
ODTK_CONV_CLUSTER_RES=1 $LB --tag r2_clres1 --only "+res" > gpurun_out/run2_lb_clres1.log 2>&1
ODTK_CONV_CLUSTER_RES=2 $LB --tag r2_clres2 --only "+res" > gpurun_out/run2_lb_clres2.log 2>&1
ODTK_CONV_CLUSTER_1X1=1 $LB --tag r2_cl1x1 --only conv1x1 > gpurun_out/run2_lb_cl1x1.log 2>&1
ODTK_CONV_BN_SHRINK=0 $LB --tag r2_shrink0 --only "256->256" > gpurun_out/run2_lb_shrink0.log 2>&1
ODTK_CONV_S2_BOX=2 $LB --tag r2_s2box2 --only " s2" > gpurun_out/run2_lb_s2box2.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/run2_bench.json 2> gpurun_out/run2_bench.err
tail -c 400 gpurun_out/run2_bench.json; tail -5 gpurun_out/run2_bench.err
