#!/bin/bash
# tools/gpu_round.sh -- what one gpurun call does: GPU parity tests, smoke, bench; logs under gpurun_out/
mkdir -p gpurun_out
export QTORCH_QUIET=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 --durations=5 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py 2>&1 | tail -2 | tee gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.log
