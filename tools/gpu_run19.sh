#!/bin/bash
mkdir -p gpurun_out
for cfg in "6 4" "5 4" "3 4" "3 8" "5 2" "4 2" "4 6"; do
  set -- $cfg
  echo "NR=$1 NW=$2"
  ODTK_BNECK_NR=$1 ODTK_BNECK_NW=$2 timeout 300 python tools/layer_bench.py --reps 10 --tag r19_$1_$2 --only bneck 2>&1 | grep bneck | cut -c1-120
done
