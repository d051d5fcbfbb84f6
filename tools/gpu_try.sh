#!/bin/bash
mkdir -p gpurun_out
export QTORCH_QUIET=1
for shape in "10 10 3 7 8 9 0 1 2" "6 14 3 1 3 5 0 6 12" "9 11 3 0 4 8 2 5 9" "10 10 3 7 8 9 0 1 2"; do
timeout 120 python tools/prof_step.py $shape 5 2>&1 | tail -3 | tee -a gpurun_out/try.log
done
timeout 600 python bench.py --steps 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_try.log
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_try.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['plan_launches_per_term'], d['kernel_time_ms_by_kind'], 'roof', d['roofline']['achieved'], 'sliced', d['sliced']['ms_per_amplitude'], d['sliced']['matches_reference_1e-10'], 'minfill', d['minfill_plan']['ms_per_term'], d['minfill_plan']['matches_reference_1e-10'])
PY
