#!/bin/bash
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 -k "cli or minfill" 2>&1 | tail -6 | tee gpurun_out/try.log
timeout 600 python bench.py --steps 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_try.log
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_try.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'sliced', d['sliced']['ms_per_amplitude'], d['minfill_plan'])
PY
