#!/bin/bash
# tools/gpu_try.sh -- scratch: whatever is being tried on the GPU box right now
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 300 python -m pytest tests/test_gpu_steps.py tests/test_maxcut.py tests/test_gpu_networks.py -x -q -m gpu --timeout 200 -k "micro or maxcut or qaoa or clusters or ghz or qft" 2>&1 | tail -2
timeout 100 python tools/prof_maxcut.py 2>&1 | tail -2 | cut -c1-20,100-220
