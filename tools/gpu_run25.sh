#!/bin/bash
mkdir -p gpurun_out
for tp in 0 1 2 3; do
echo "TWOPASS=$tp"
ODTK_CONV_CAND_TWOPASS=$tp timeout 120 python tools/layer_bench.py --calibrated --reps 3 --tag r25_$tp --only cand > gpurun_out/run25_$tp.log 2>&1
grep -E "cand|Error|error" gpurun_out/run25_$tp.log | cut -c1-140 | head -8; tail -3 gpurun_out/run25_$tp.log | cut -c1-200
done
