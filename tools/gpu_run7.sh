#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/run7_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/run7_pytest.log
grep -E "AssertionError|passed|failed|^FAILED|rc=" gpurun_out/run7_pytest.log | head
ODTK_FILTER_BULK=1 timeout 600 python -m pytest tests/test_gpu_postproc.py -q > gpurun_out/run7_pytest_bulk.log 2>&1; echo "rc=$?" >> gpurun_out/run7_pytest_bulk.log
tail -3 gpurun_out/run7_pytest_bulk.log
for b in 0 1 2 4 6; do
  ODTK_FILTER_BULK=$b timeout 300 python bench.py --config postproc --steps 50 --no-cpu-baseline --no-e2e > gpurun_out/run7_postproc_bulk$b.json 2>> gpurun_out/run7_postproc.err
  python -c "
import json; d=json.loads(open('gpurun_out/run7_postproc_bulk$b.json').read().strip().splitlines()[-1]); print('bulk$b', d['us_per_image'], d['roofline']['achieved'])"
done
timeout 300 python tools/layer_bench.py --tag r7_1x1 --only conv1x1 > gpurun_out/run7_lb_1x1.log 2>&1
ODTK_CONV_DEEP_1X1=0 timeout 300 python tools/layer_bench.py --tag r7_1x1_nodeep --only conv1x1 > gpurun_out/run7_lb_1x1_nodeep.log 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-postproc > gpurun_out/run7_bench.json 2> gpurun_out/run7_bench.err
tail -c 300 gpurun_out/run7_bench.json
