#!/bin/bash
# tools/gpu_try.sh -- scratch: whatever is being tried on the GPU box right now
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 600 python -m pytest tests/test_gpu_steps.py -x -q -k "shared_sum or tma_fed or gett" 2>&1 | tail -5
