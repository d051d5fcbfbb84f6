#!/bin/bash
# tools/gpu_try.sh -- scratch: whatever is being tried on the GPU box right now
mkdir -p gpurun_out
export QTORCH_QUIET=1
QTB_MICRO_CLUSTER=4 timeout 420 python -m pytest tests -x -q -m gpu --timeout 200 --deselect tests/test_gpu_networks.py::test_reference_test_suite_drop_in 2>&1 | tail -5 | tee gpurun_out/pytest_cluster4.log
