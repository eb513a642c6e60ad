#!/bin/bash
# usage: bash tools/gpu_scale.sh N [extra bench args...]   (run under gpurun --gpus N)
N=$1; shift
mkdir -p gpurun_out
set -x
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_gather.py -q > gpurun_out/scale_gather_test.log 2>&1; echo "rc=$?" >> gpurun_out/scale_gather_test.log
  tail -5 gpurun_out/scale_gather_test.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 "$@" > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
tail -c 400 gpurun_out/scale_n$N.json; tail -3 gpurun_out/scale_n$N.err
if [ "$N" = "8" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config rn101x8 --steps 30 --warmup 5 > gpurun_out/scale_rn101x8.json 2> gpurun_out/scale_rn101x8.err
  tail -c 400 gpurun_out/scale_rn101x8.json; tail -3 gpurun_out/scale_rn101x8.err
  ODTK_BENCH_GATHER=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 30 --warmup 5 --no-e2e > gpurun_out/scale_n8_nccl.json 2> gpurun_out/scale_n8_nccl.err
  tail -c 300 gpurun_out/scale_n8_nccl.json
fi
