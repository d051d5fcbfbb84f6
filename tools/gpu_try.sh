#!/bin/bash
# tools/gpu_try.sh -- scratch: a quick GPU check between full rounds
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 900 python -m pytest tests/test_maxcut.py tests/test_gpu_cli.py -x -q -m gpu --timeout 900 2>&1 | tail -15 > gpurun_out/pytest_try.log
tail -15 gpurun_out/pytest_try.log
timeout 600 python tools/prof_maxcut.py > gpurun_out/prof_maxcut.log 2>&1
tail -5 gpurun_out/prof_maxcut.log
timeout 900 python tools/prof_sliced2.py --shares > gpurun_out/prof_sliced2_shares.log 2>&1
tail -30 gpurun_out/prof_sliced2_shares.log
