#!/bin/bash
# tools/gpu_try.sh -- scratch: whatever is being tried on the GPU box right now
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 500 python -m pytest tests/test_gpu_cli.py -x -q -m gpu --timeout 300 -k "shim" 2>&1 | tail -8
