#!/bin/bash
# quick A/B: GPU tests, per-kernel times (1024 x 60 s one-warp and pipelined, 8192 x 30 s), resident step time
mkdir -p gpurun_out; O=gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python profiles/tools/kernel_times.py 1024 60
SPEEDY_K4_PIPELINE=1 python profiles/tools/kernel_times.py 1024 60
python profiles/tools/kernel_times.py 8192 30 16000 1 3.5
python profiles/tools/step_time.py 2>&1 | tail -2
