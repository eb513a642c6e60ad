#!/bin/bash
# round-2 GPU run 10: stem output geometry fix; full gpu suite; per-launch ncu table of one step; bench
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/run10_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/run10_pytest.log
grep -E "passed|failed|^FAILED|^ERROR|rc=" gpurun_out/run10_pytest.log | head -40
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file gpurun_out/run10_step_launches.csv python tools/capture_step.py --trace gpurun_out/run10_step_trace.json > gpurun_out/run10_capture.log 2>&1
python tools/layer_table.py gpurun_out/run10_step_launches.csv gpurun_out/run10_step_trace.json --out gpurun_out/run10_layer_table > gpurun_out/run10_layer_table.txt 2>&1
head -50 gpurun_out/run10_layer_table.txt
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/run10_bench.json 2> gpurun_out/run10_bench.err
tail -c 600 gpurun_out/run10_bench.json; tail -5 gpurun_out/run10_bench.err
