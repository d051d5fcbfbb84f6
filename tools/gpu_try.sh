#!/bin/bash
# tools/gpu_try.sh -- scratch: a quick GPU check between full rounds
mkdir -p gpurun_out
export QTORCH_QUIET=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 --durations=8 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/prof_sliced2.py > gpurun_out/prof_sliced2.log 2>&1
tail -50 gpurun_out/prof_sliced2.log
