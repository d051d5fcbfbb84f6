#!/bin/bash
# tools/gpu_try.sh -- scratch: whatever is being tried on the GPU box right now
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 500 python -m pytest tests/test_gpu_steps.py tests/test_gpu_networks.py tests/test_slicing.py -x -q -m gpu --timeout 200 --deselect tests/test_gpu_networks.py::test_reference_test_suite_drop_in 2>&1 | tail -3
for c in 1 0; do if [ $c = 1 ]; then export QTB_MICRO_CLUSTER=1; else unset QTB_MICRO_CLUSTER; fi; echo "== QTB_MICRO_CLUSTER=${QTB_MICRO_CLUSTER:-auto}"; timeout 100 python tools/prof_micro.py qaoa30_z27z29 3 | tail -1; timeout 100 python tools/prof_micro.py ghz1000_zeros 3 | tail -1; timeout 100 python tools/prof_micro.py rand42_cn4_d20_zeros 3 | tail -1; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_try.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_try.log').read())
print('value %.2f e2e %.2f ms %.3f kinds %s'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['kernel_time_ms_by_kind']))
print('sliced %.1f %.3f | cfg2 %.1f %.3f'%(d['sliced']['value'], d['sliced']['ms_per_amplitude'], d['sliced_cfg2']['value'], d['sliced_cfg2']['ms_per_amplitude']))
PY
