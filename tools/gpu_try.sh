#!/bin/bash
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 900 python -m pytest tests/test_gpu_steps.py tests/test_gpu_networks.py tests/test_slicing.py tests/test_maxcut.py -x -q -m gpu --timeout 600 -k "not drop_in" 2>&1 | tail -5 | tee gpurun_out/try.log
for v in "QTB_NO_APPLY=0"; do
echo "== $v" | tee -a gpurun_out/try.log
for shape in "13 1 1 12 0" "13 1 1 0 0" "1 13 1 0 4" "14 2 1 13 0" "14 2 1 0 1" "2 14 1 0 7" "2 14 1 1 0" "14 2 2 0 1 1 0" "14 2 2 5 11 0 1" "13 3 2 0 12 2 0" "3 13 2 0 2 12 3"; do
env $v timeout 120 python tools/prof_step.py $shape 3 2>&1 | tail -1 | tee -a gpurun_out/try.log
done
done
timeout 600 python bench.py --steps 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_try.log
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_try.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['plan_launches_per_term'], d['kernel_time_ms_by_kind'], 'roof', d['roofline']['achieved'], 'sliced', d['sliced']['ms_per_amplitude'], d['sliced']['matches_reference_1e-10'], 'minfill', d['minfill_plan']['ms_per_term'], d['minfill_plan']['matches_reference_1e-10'])
PY
