#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err
tail -c 500 gpurun_out/scale_n8.json; tail -3 gpurun_out/scale_n8.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config rn101x8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_rn101x8.json 2> gpurun_out/scale_rn101x8.err
tail -c 300 gpurun_out/scale_rn101x8.json; tail -3 gpurun_out/scale_rn101x8.err | cut -c1-300
