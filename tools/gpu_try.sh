#!/bin/bash
# tools/gpu_try.sh -- scratch: whatever is being tried on the GPU box right now
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_networks.py::test_reference_test_suite_drop_in 2>&1 | tail -4
timeout 400 python tools/prof_maxcut.py --per-term > gpurun_out/maxcut.txt 2>&1; cat gpurun_out/maxcut.txt
timeout 100 python tools/prof_micro.py ghz1000_zeros 3 | tail -1
timeout 100 python tools/prof_micro.py qaoa30_z27z29 3 | tail -1
