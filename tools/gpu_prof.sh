#!/bin/bash
# tools/gpu_prof.sh -- ncu evidence: launch list of a short bench run + full captures of the dominant kernels
mkdir -p gpurun_out
export QTORCH_QUIET=1
NCU=/usr/local/cuda/bin/ncu
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# (10,10,k=3 -> 14): the config-2 big x big step;  (14,14,k=14 -> 0): the closing inner product; (14,2,k=1 -> 14): streaming gate application
timeout 300 $NCU --set full --clock-control none --import-source on -k regex:k_gett -s 1 -c 1 -o gpurun_out/prof_gett python tools/prof_step.py 10 10 3 0 2 5 1 7 6 2 > gpurun_out/prof_gett.log 2>&1
timeout 300 $NCU --set full --clock-control none --import-source on -k regex:k_dot -s 1 -c 1 -o gpurun_out/prof_dot python tools/prof_step.py 14 14 14 0 1 2 3 4 5 6 7 8 9 10 11 12 13 9 6 8 7 0 5 13 2 10 12 4 1 3 11 2 > gpurun_out/prof_dot.log 2>&1
timeout 300 $NCU --set full --clock-control none --import-source on -k regex:k_apply -s 1 -c 1 -o gpurun_out/prof_stream python tools/prof_step.py 14 2 1 13 0 2 > gpurun_out/prof_stream.log 2>&1

timeout 300 $NCU --set full --clock-control none --import-source on -k regex:k_gett -s 1 -c 1 -o gpurun_out/prof_fused python tools/prof_fused.py 2 > gpurun_out/prof_fused.log 2>&1
ls -la gpurun_out
