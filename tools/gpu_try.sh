#!/bin/bash
mkdir -p gpurun_out
export QTORCH_QUIET=1
timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_maxcut.py tests/test_slicing.py -x -q -m gpu --timeout 600 2>&1 | tail -5 | tee gpurun_out/try.log
cd /tmp && for n in 1 2; do
  python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n $GRAFT_REPO_ROOT/qtorch_b200/bin/maxcutQAOA $GRAFT_REPO_ROOT/tests/golden/Samples/3regRand30Node50.dgf 2 0 /tmp/angles$n.txt 60 2>&1 | grep -v "^$" | tail -4 | tee -a $GRAFT_REPO_ROOT/gpurun_out/try.log
done
