#!/bin/bash
# tools/gpu_try.sh -- scratch: a quick GPU check between full rounds (parity tests without the 90 s drop-in suite + a short bench)
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 -k "not drop_in" 2>&1 | tail -5 | tee gpurun_out/try.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_try.log
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_try.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['config']['plan_launches_per_term'], d['kernel_time_ms_by_kind'], 'roof', d['roofline']['achieved'], 'sliced', d['sliced']['ms_per_amplitude'], d['sliced']['matches_reference_1e-10'], 'minfill', d['minfill_plan']['ms_per_term'], d['minfill_plan']['matches_reference_1e-10'])
PY
