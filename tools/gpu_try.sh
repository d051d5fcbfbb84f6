#!/bin/bash
# tools/gpu_try.sh -- scratch: whatever is being tried on the GPU box right now
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 900 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 100 python tools/prof_maxcut.py 2>&1 | tail -2 | tee gpurun_out/maxcut.txt
timeout 100 python tools/prof_micro.py ghz1000_zeros 3 | tail -1
