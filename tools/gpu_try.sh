#!/bin/bash
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 600 python -m pytest tests/test_slicing.py tests/test_gpu_networks.py -x -q -m gpu --timeout 300 -k "slic or run_slots or jobs_in_flight or value_matches" 2>&1 | tail -8 | tee gpurun_out/try.log
timeout 300 python tools/prof_e2e.py 2>&1 | tee gpurun_out/prof_e2e.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_try.log
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_try.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['plan_launches_per_term'], d['kernel_time_ms_by_kind'], 'roof', d['roofline']['achieved'], 'sliced', d['sliced']['ms_per_amplitude'], d['sliced']['matches_reference_1e-10'], 'minfill', d['minfill_plan']['ms_per_term'], d['minfill_plan']['matches_reference_1e-10'])
PY
