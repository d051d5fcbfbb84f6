#!/bin/bash
# tools/gpu_multi.sh N -- bench.py under torchrun on N GPUs of one box (the driver's launch line)
N=${1:-2}
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err
tail -c 6000 gpurun_out/bench_${N}gpu.log; tail -5 gpurun_out/bench_${N}gpu.err
