#!/bin/bash
mkdir -p gpurun_out
export QTORCH_QUIET=1
for L in 5 6 7; do
echo "== QTB_MICRO_LOG4=$L" | tee -a gpurun_out/try.log
QTB_MICRO_LOG4=$L timeout 600 python bench.py --steps 10 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_m$L.log
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_m$L.log').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['config']['plan_launches_per_term'], d['kernel_time_ms_by_kind'], 'sliced ms', d['sliced']['ms_per_amplitude'], d['sliced']['matches_reference_1e-10'])
PY
done
timeout 900 python -m pytest tests -x -q -m gpu --timeout 600 -k "not drop_in" 2>&1 | tail -3 | tee -a gpurun_out/try.log
timeout 500 python tools/bench_configs.py 2>&1 | grep "^{" | tee gpurun_out/configs.jsonl | cut -c1-250
