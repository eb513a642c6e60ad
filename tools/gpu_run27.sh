#!/bin/bash
# round-2 evidence run: full gpu suite, every BASELINE config, ncu launch lists / layer table / full captures, sanitizer
set -x
O=gpurun_out/r02
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "rc=$?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log
timeout 900 python bench.py --steps 30 --warmup 5 > $O/bench_full.json 2> $O/bench_full.err
tail -c 400 $O/bench_full.json
for cfg in b1 b8 rotated postproc postproc_rotated rn101x8; do
  timeout 600 python bench.py --config $cfg --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_$cfg.json 2> $O/bench_$cfg.err
  tail -c 200 $O/bench_$cfg.json; echo
done
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
tail -c 300 $O/bench_reference.json
# ncu: launch list of the bench command itself, and the per-launch table of one eager step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-postproc > $O/bench_under_ncu.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file $O/step_launches.csv python tools/capture_step.py --trace $O/step_trace.json > $O/capture.log 2>&1
python tools/layer_table.py $O/step_launches.csv $O/step_trace.json --out $O/layer_table > $O/layer_table.txt 2>&1
head -12 $O/layer_table.txt
# ncu full captures: the fused bottleneck kernel (C1 = 64), the head tower launch, the class-head final launch
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:bottleneck_tail -s 1 -c 1 -f -o $O/ncu_bneck_l1 python tools/capture_step.py > $O/ncu_bneck.log 2>&1

# compute-sanitizer on the conv / bottleneck cases
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_cases.py conv > $O/sanitizer_memcheck_conv.log 2>&1
tail -3 $O/sanitizer_memcheck_conv.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_cases.py conv > $O/sanitizer_racecheck_conv.log 2>&1
tail -3 $O/sanitizer_racecheck_conv.log
ls -la $O
