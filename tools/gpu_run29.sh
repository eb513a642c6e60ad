#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -q -k "depthwise or relu6" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_model.py -q -k "MobileNet" 2>&1 | tail -15 | cut -c1-250
timeout 300 python bench.py --backbone MobileNetV2FPN --batch 8 --steps 20 --warmup 3 --no-cpu-baseline --no-postproc 2>&1 | tail -c 700
