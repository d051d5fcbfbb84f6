#!/bin/bash
# tools/gpu_try.sh -- scratch: 8-GPU bench
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --no-cpu-baseline > gpurun_out/bench_8gpu_try.log 2>gpurun_out/bench_8gpu_try.err
tail -5 gpurun_out/bench_8gpu_try.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_8gpu_try.log').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e'])
print('sliced', {k:d['sliced'][k] for k in ('value','ms_per_amplitude','wall_ms_per_amplitude','e2e','slices','matches_reference_1e-10')})
print('sliced_cfg2', {k:d['sliced_cfg2'][k] for k in ('value','ms_per_amplitude','e2e','matches_reference_1e-10')})
print('maxcut', d['maxcut']['p1'], d['maxcut']['p2'])
PY
