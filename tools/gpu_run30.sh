#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_conv.py -q -x -k "bottleneck" 2>&1 | tail -12 | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_model.py -q -x 2>&1 | tail -4 | cut -c1-200
timeout 200 python tools/layer_bench.py --tag r30 --only "bneck" 2>&1 | grep bneck | cut -c1-120
ODTK_FUSED_CONV1=0 timeout 200 python tools/layer_bench.py --tag r30_off --only "200x320" 2>&1 | grep -E "bneck|256->64|256->128" | cut -c1-120
for v in 0 1; do
ODTK_FUSED_CONV1=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fused_conv1=$v', d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
