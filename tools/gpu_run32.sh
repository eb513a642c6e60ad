#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for cfg in postproc postproc_rotated; do
timeout 300 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/run32_$cfg.json 2> gpurun_out/run32_$cfg.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/run32_$cfg.json").read().strip().splitlines()[-1]); print("$cfg", d["value"], "img/s", d["us_per_image"], "us/img", d["roofline"], d["e2e"])
except Exception as e: print("$cfg failed", e, open("gpurun_out/run32_$cfg.err").read()[-800:])
PY
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('full', d['value'], d['ms_per_step'], json.dumps(d['postproc'])[:900])"
