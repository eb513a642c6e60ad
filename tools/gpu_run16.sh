#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bottleneck_tail -s 2 -c 1 -f -o gpurun_out/run16_bt_l1 python tools/bt_run.py > gpurun_out/run16_ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bottleneck_tail -s 5 -c 1 -f -o gpurun_out/run16_bt_l2 python tools/bt_run.py > gpurun_out/run16_ncu2.log 2>&1
tail -3 gpurun_out/run16_ncu1.log
ls -la gpurun_out/*.ncu-rep
