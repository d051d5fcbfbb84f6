#!/bin/bash
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 600 python -m pytest tests -x -q -m gpu --timeout 300 -k "maxcut or qaoa" 2>&1 | tail -15 | tee gpurun_out/try.log
timeout 600 python tools/bench_configs.py 2>&1 | tail -12 | tee gpurun_out/configs.jsonl
