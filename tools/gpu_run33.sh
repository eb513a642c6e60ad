#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/run33_bench.json 2> gpurun_out/run33_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/run33_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["postproc"]["us_per_image"], d["postproc"]["us_per_image_one_batch_at_a_time"], d["cpu_baseline"]["value"], d["clocks"])
PY
