#!/bin/bash
mkdir -p gpurun_out
for cfg in "0 0" "84 64" "74 74" "96 52" "64 84" "104 44"; do
  set -- $cfg
  timeout 300 python tools/pipe_bench.py $1 $2 20 2>&1 | tail -1
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'])"
