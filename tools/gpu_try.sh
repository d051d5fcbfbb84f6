#!/bin/bash
# tools/gpu_try.sh -- scratch: whatever is being tried on the GPU box right now
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 500 python -m pytest tests/test_gpu_steps.py tests/test_gpu_networks.py tests/test_slicing.py -x -q -m gpu --timeout 200 --deselect tests/test_gpu_networks.py::test_reference_test_suite_drop_in 2>&1 | tail -3
for v in 0 1; do
QTB_PDL=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_try.log; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_try.log').read())
print('PDL=$v value %.2f e2e %.2f ms %.3f gett %.4f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['avg_ms']))
print('   sliced %.1f %.3f | cfg2 %.1f %.3f | minfill %.3f'%(d['sliced']['value'], d['sliced']['ms_per_amplitude'], d['sliced_cfg2']['value'], d['sliced_cfg2']['ms_per_amplitude'], d['minfill_plan']['ms_per_term']))
PY
done
