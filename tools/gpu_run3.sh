#!/bin/bash
# round-2 GPU run 3: residual through the pipeline, narrow cta_group::2, iou / loss kernels; stem ncu; clocks under load
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/run3_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/run3_pytest.log
tail -15 gpurun_out/run3_pytest.log
LB="timeout 300 python tools/layer_bench.py"
$LB --tag r3base > gpurun_out/run3_lb_base.log 2>&1
ODTK_CONV_RES_PIPE=0 $LB --tag r3_respipe0 --only "+res" > gpurun_out/run3_lb_respipe0.log 2>&1
ODTK_CONV_TWO_NARROW=0 $LB --tag r3_twonarrow0 --only "f32" > gpurun_out/run3_lb_twonarrow0.log 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,temperature.gpu --format=csv -lms 20 > gpurun_out/run3_clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/run3_bench.json 2> gpurun_out/run3_bench.err
kill $SMI
tail -c 300 gpurun_out/run3_bench.json; tail -5 gpurun_out/run3_bench.err
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:stem_pool -c 1 -f -o gpurun_out/run3_ncu_stempool python tools/capture_step.py > gpurun_out/run3_ncu_stempool.log 2>&1
for spec in "26 res1024" "69 boxfinal" "2 conv64"; do
  set -- $spec
  timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:conv_gemm -s $1 -c 1 -f -o gpurun_out/run3_ncu_$2 python tools/capture_step.py > gpurun_out/run3_ncu_$2.log 2>&1
done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/run3_step_launches.csv python tools/capture_step.py --trace gpurun_out/run3_step_trace.json > gpurun_out/run3_capture.log 2>&1
ls gpurun_out | grep run3
