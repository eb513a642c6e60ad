#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -q -x -k "bottleneck" > gpurun_out/run18_bt.log 2>&1; echo "rc=$?" >> gpurun_out/run18_bt.log
tail -3 gpurun_out/run18_bt.log | cut -c1-200
timeout 300 python tools/layer_bench.py --tag r18_base --only bneck > gpurun_out/run18_lb_base.log 2>&1
ODTK_BNECK_NR=2 timeout 300 python tools/layer_bench.py --tag r18_nr2 --only bneck > gpurun_out/run18_lb_nr2.log 2>&1
ODTK_BNECK_NR=3 timeout 300 python tools/layer_bench.py --tag r18_nr3 --only bneck > gpurun_out/run18_lb_nr3.log 2>&1
grep -h bneck gpurun_out/run18_lb_*.log | cut -c1-120
ODTK_B200_LIB=tools/_ab/lib_btprof.so timeout 300 python tools/bt_prof.py > gpurun_out/run18_prof.log 2>&1
grep -E "shape|mma|prod|epi" gpurun_out/run18_prof.log
