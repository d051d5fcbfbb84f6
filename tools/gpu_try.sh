#!/bin/bash
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 100 python tools/check_tma.py 2>&1 | tail -9
for shape in "10 10 3 0 2 5 1 7 6" "6 14 3 0 2 3 2 7 9"; do
  echo "-- $shape TMA"; timeout 60 python tools/prof_step.py $shape 3 | tail -1
done
