#!/bin/bash
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 900 python -m pytest tests/test_maxcut.py tests/test_gpu_steps.py tests/test_gpu_networks.py -x -q -m gpu --timeout 600 -k "not drop_in" 2>&1 | tail -5 | tee gpurun_out/try.log
timeout 600 python tools/bench_configs.py 2>&1 | grep maxcut | cut -c1-400 | tee gpurun_out/configs_try.jsonl
