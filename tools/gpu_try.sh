#!/bin/bash
# tools/gpu_try.sh -- scratch: whatever is being tried on the GPU box right now
mkdir -p gpurun_out
export QTORCH_QUIET=1
NCU=/usr/local/cuda/bin/ncu
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches.csv
