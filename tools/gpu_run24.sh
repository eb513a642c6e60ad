#!/bin/bash
mkdir -p gpurun_out
for tp in 1 0; do
ODTK_CONV_CAND_TWOPASS=$tp ODTK_BENCH_INSTEP=gpurun_out/run24_instep_$tp.json timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-postproc --no-e2e > gpurun_out/run24_bench_$tp.json 2>> gpurun_out/run24_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/run24_bench_$tp.json").read().strip().splitlines()[-1]); print("twopass=$tp", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"])
d=json.load(open('gpurun_out/run24_instep_$tp.json'))
for r in d['rows']:
    if '720' in r['layer'] or 'stem' in r['layer'] or '215' in r['layer']: print("%-46s n=%-2d %8.1f us %5.1f%% %7.1f TF"%(r['layer'][:46],r['n'],r['us'],100*r['share'],r['tflops']))
PY
done
