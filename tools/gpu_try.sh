#!/bin/bash
mkdir -p gpurun_out
export QTORCH_QUIET=1
echo "== C1 8 math warps + producer warpgroup" | tee gpurun_out/try.log
timeout 60 python tools/prof_step.py 10 10 3 0 2 5 1 7 6 3 | tee -a gpurun_out/try.log
timeout 60 python tools/prof_step.py 6 14 3 0 2 3 2 7 9 3 | tee -a gpurun_out/try.log
echo "== C1 16 math warps" | tee -a gpurun_out/try.log
QTB_GETT_C1=1 timeout 60 python tools/prof_step.py 10 10 3 0 2 5 1 7 6 3 | tee -a gpurun_out/try.log
QTB_GETT_C1=1 timeout 60 python tools/prof_step.py 9 11 3 0 4 6 6 8 5 3 | tee -a gpurun_out/try.log
echo "== streaming" | tee -a gpurun_out/try.log
timeout 60 python tools/prof_step.py 14 2 1 13 0 3 | tee -a gpurun_out/try.log
timeout 60 python tools/prof_step.py 14 4 2 3 9 0 1 3 | tee -a gpurun_out/try.log
timeout 60 python tools/prof_step.py 13 1 1 5 0 3 | tee -a gpurun_out/try.log
echo "== dot" | tee -a gpurun_out/try.log
timeout 60 python tools/prof_step.py 14 14 14 0 1 2 3 4 5 6 7 8 9 10 11 12 13 9 6 8 7 0 5 13 2 10 12 4 1 3 11 3 | tee -a gpurun_out/try.log
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -3 | tee -a gpurun_out/try.log
QTB_GETT_C1=1 timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 -k "gett or config2 or plan_api" 2>&1 | tail -3 | tee -a gpurun_out/try.log
