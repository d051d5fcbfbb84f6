#!/bin/bash
mkdir -p gpurun_out
export QTORCH_QUIET=1
P="timeout 90 python tools/prof_step.py"
echo "== C1 (default 16 math warps)" | tee gpurun_out/try.log
$P 10 10 3 0 2 5 1 7 6 3 | tee -a gpurun_out/try.log
$P 6 14 3 0 2 3 2 7 9 3 | tee -a gpurun_out/try.log
echo "== streaming" | tee -a gpurun_out/try.log
$P 14 2 1 13 0 3 | tee -a gpurun_out/try.log
$P 14 4 2 3 9 0 1 3 | tee -a gpurun_out/try.log
$P 13 1 1 5 0 3 | tee -a gpurun_out/try.log
$P 13 3 1 2 1 3 | tee -a gpurun_out/try.log
$P 4 13 2 0 3 12 0 3 | tee -a gpurun_out/try.log
echo "== dot" | tee -a gpurun_out/try.log
$P 14 14 14 0 1 2 3 4 5 6 7 8 9 10 11 12 13 9 6 8 7 0 5 13 2 10 12 4 1 3 11 3 | tee -a gpurun_out/try.log
timeout 600 python -m pytest tests -x -q -m gpu --timeout 300 2>&1 | tail -3 | tee -a gpurun_out/try.log
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench.log
