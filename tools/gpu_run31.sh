#!/bin/bash
mkdir -p gpurun_out
for late in 0 1; do
echo "G3_LATE=$late"
ODTK_BNECK_G3_LATE=$late timeout 120 python -m pytest tests/test_gpu_conv.py -q -x -k "next_conv1" 2>&1 | tail -1
ODTK_BNECK_G3_LATE=$late timeout 200 python tools/layer_bench.py --tag r31_$late --only "bneck" 2>&1 | grep "+up" | cut -c1-120
done
ODTK_BNECK_G3_LATE=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('late=1', d['value'], d['ms_per_step'], d['roofline']['frac'])"
ODTK_BNECK_G3_LATE=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('late=0', d['value'], d['ms_per_step'], d['roofline']['frac'])"
