#!/bin/bash
mkdir -p gpurun_out
export QTORCH_QUIET=1
P="timeout 90 python tools/prof_step.py"
echo "== 3M 16 warps pipelined" | tee gpurun_out/try.log
QTB_GETT_C1=2 $P 10 10 3 0 2 5 1 7 6 3 | tail -1 | tee -a gpurun_out/try.log
QTB_GETT_C1=2 $P 6 14 3 0 2 3 2 7 9 3 | tail -1 | tee -a gpurun_out/try.log
echo "== 3M 8 fat warps pipelined" | tee -a gpurun_out/try.log
QTB_GETT_C1=3 $P 10 10 3 0 2 5 1 7 6 3 | tail -1 | tee -a gpurun_out/try.log
QTB_GETT_C1=3 $P 6 14 3 0 2 3 2 7 9 3 | tail -1 | tee -a gpurun_out/try.log
QTB_GETT_C1=3 $P 9 11 3 0 4 6 6 8 5 3 | tail -1 | tee -a gpurun_out/try.log
QTB_GETT_C1=3 $P 9 9 2 0 4 6 8 3 | tail -1 | tee -a gpurun_out/try.log
QTB_GETT_C1=2 timeout 600 python -m pytest tests -x -q -m gpu --timeout 300 -k "gett or config2 or plan_api or linearity" 2>&1 | tail -2 | tee -a gpurun_out/try.log
QTB_GETT_C1=3 timeout 600 python -m pytest tests -x -q -m gpu --timeout 300 -k "gett or config2 or plan_api or linearity" 2>&1 | tail -2 | tee -a gpurun_out/try.log
