#!/bin/bash
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 900 python -m pytest tests/test_gpu_networks.py -x -q -m gpu --timeout 600 -k "cost" 2>&1 | tail -3 | tee gpurun_out/try.log
: > gpurun_out/maxcut_multi.jsonl
for p in 1 2; do
  timeout 300 python tools/bench_maxcut_multi.py --p $p 2>&1 | tail -1 | tee -a gpurun_out/maxcut_multi.jsonl
  for n in 2 4; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n tools/bench_maxcut_multi.py --p $p 2>&1 | tail -1 | tee -a gpurun_out/maxcut_multi.jsonl
  done
done
