#!/bin/bash
# tools/gpu_try.sh -- scratch: whatever is being tried on the GPU box right now
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 400 python -m pytest tests/test_gpu_steps.py tests/test_maxcut.py tests/test_gpu_networks.py -x -q -m gpu --timeout 200 --deselect tests/test_gpu_networks.py::test_reference_test_suite_drop_in 2>&1 | tail -3
timeout 100 python tools/micro_timeline.py 18 2>&1 | tail -17 | tee gpurun_out/micro_timeline2.txt
timeout 100 python tools/prof_maxcut.py 2>&1 | tail -2 | tee gpurun_out/maxcut.txt
timeout 100 python tools/prof_micro.py ghz1000_zeros 3 | tail -1
