#!/bin/bash
# tools/gpu_try.sh -- scratch: whatever is being tried on the GPU box right now
mkdir -p gpurun_out
export QTORCH_QUIET=1
CS=/usr/local/cuda/bin/compute-sanitizer
( echo "== memcheck, default kernels"; timeout 400 $CS --tool memcheck --print-limit 5 python tools/sanitize_small.py 2>&1 | tail -16
  echo "== memcheck, QTB_GETT_C1=5 QTB_MICRO_CLUSTER=4"; QTB_GETT_C1=5 QTB_MICRO_CLUSTER=4 timeout 400 $CS --tool memcheck --print-limit 5 python tools/sanitize_small.py 2>&1 | tail -16 ) | tee gpurun_out/sanitizer.txt
