#!/bin/bash
# tools/gpu_try.sh -- scratch: whatever is being tried on the GPU box right now
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 400 python -m pytest tests/test_maxcut.py tests/test_gpu_networks.py -x -q -m gpu --timeout 200 -k "maxcut or qaoa or clusters or config2" 2>&1 | tail -3
timeout 100 python tools/prof_maxcut.py 2>&1 | tail -2
timeout 100 python tools/prof_micro.py qaoa30_z27z29 3 | tail -1
