#!/bin/bash
# timing experiments with deliberately wrong maths (tools/_build/lib_*.so); never a bench value
mkdir -p gpurun_out
export QTORCH_QUIET=1
run() { QTB_PRESUM=0 timeout 120 python tools/prof_step.py 10 10 3 7 8 9 0 1 2 3 2>&1 | tail -1; QTB_PRESUM=0 timeout 120 python tools/prof_step.py 6 14 3 1 3 5 0 6 12 3 2>&1 | tail -1; }
echo "== product"; run
cp qtorch_b200/libqtorch_b200.so /tmp/lib_product.so
for v in "$@"; do
  echo "== $v"; cp tools/_build/lib_$v.so qtorch_b200/libqtorch_b200.so; run
done
cp /tmp/lib_product.so qtorch_b200/libqtorch_b200.so
