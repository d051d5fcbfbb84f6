#!/bin/bash
# tools/gpu_try.sh -- scratch: whatever is being tried on the GPU box right now
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 400 python -m pytest tests/test_gpu_steps.py tests/test_maxcut.py -x -q -m gpu --timeout 200 -k "programmatic or maxcut or qaoa" 2>&1 | tail -3
