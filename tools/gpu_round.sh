#!/bin/bash
# tools/gpu_round.sh -- what one gpurun call does: GPU parity tests, smoke, bench; logs under gpurun_out/
mkdir -p gpurun_out
export QTORCH_QUIET=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log
