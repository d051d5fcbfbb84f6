#!/bin/bash
# tools/gpu_try.sh -- scratch: whatever is being tried on the GPU box right now
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 300 python tools/micro_timeline.py 18 2>&1 | tail -40 | tee gpurun_out/micro_timeline.txt
