#!/bin/bash
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 900 python -m pytest tests/test_gpu_steps.py -x -q -m gpu --timeout 600 2>&1 | tail -5 | tee gpurun_out/try.log
for shape in "14 2 2 0 1 1 0" "14 3 2 0 1 0 2" "3 14 2 1 2 1 0"; do
timeout 120 python tools/prof_step.py $shape 3 2>&1 | tail -1 | tee -a gpurun_out/try.log
done
