"""N > 1 host logic on CPU: world_size-2 gloo job running the term dispatcher (round-robin dealing + one
sum-allreduce), with the per-term values computed by the CPU oracle from plans exported by the host mirror."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from qtorch_b200.dispatch import deal_round_robin, Dispatcher, qaoa_objective

WORKER = r'''
import json, os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch.distributed as dist
from qtorch_b200 import host_api
from qtorch_b200.dispatch import Dispatcher, qaoa_objective
from oracle import oracle as O
G = os.path.join(sys.argv[1], "tests", "golden")
dist.init_process_group("gloo")
qasm = os.path.join(G, "Samples", "test_JW.qasm")
ordering = os.path.join(G, "orderings", "testJW_XXXX.qbb.out")
terms = ["Z Z T T", "Z T Z T", "Z T T Z", "T Z Z T", "T Z T Z", "T T Z Z", "X X X X"]
def zz(u):
    m = os.path.join(sys.argv[2], "m%d_%d.txt" % (dist.get_rank(), u))
    open(m, "w").write(terms[u] + "\n")
    ranks, steps, inputs, _ = host_api.export_plan_linegraph(qasm, m, ordering, True)
    t, rk = list(inputs), list(ranks)
    for a, b, pa, pb in steps:
        t.append(O.contract(t[a], rk[a], t[b], rk[b], pa, pb)); rk.append(rk[a] + rk[b] - 2 * len(pa))
    return complex(t[-1][0])
d = Dispatcher.for_torch_distributed()
total, vec = d.map_reduce(len(terms), zz, want_vector=True)
fp = qaoa_objective(d, len(terms), zz)
if dist.get_rank() == 0:
    print("@@" + json.dumps({"total": [total.real, total.imag], "vec": [[v.real, v.imag] for v in vec], "fp": fp, "owned": d.owned(len(terms))}))
dist.destroy_process_group()
'''


def test_round_robin_dealing():
    assert deal_round_robin(45, 0, 8) == [0, 8, 16, 24, 32, 40]
    assert deal_round_robin(45, 7, 8) == [7, 15, 23, 31, 39]
    allu = sorted(u for r in range(8) for u in deal_round_robin(45, r, 8))
    assert allu == list(range(45))
    with pytest.raises(ValueError):
        deal_round_robin(4, 2, 2)


def test_single_rank_dispatcher():
    d = Dispatcher()
    assert d.map_reduce(5, lambda u: complex(u, -u)) == complex(10, -10)
    assert qaoa_objective(d, 3, lambda u: complex(0.5, 0.3)) == pytest.approx(0.75)


def test_two_rank_gloo_dispatch(built, tmp_path):
    script = os.path.join(str(tmp_path), "worker.py")
    open(script, "w").write(WORKER)
    env = dict(os.environ)
    env["QTORCH_QUIET"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", script, ROOT, str(tmp_path)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("@@")][0]
    res = json.loads(line[2:])
    vec = np.array([complex(*v) for v in res["vec"]])
    nets = json.load(open(os.path.join(GOLDEN, "networks.json")))
    assert abs(vec[6] - complex(*nets["testJW_XXXX"]["value"])) < 1e-12          # the X X X X term is a golden value
    assert abs(complex(*res["total"]) - vec.sum()) < 1e-12
    assert abs(res["fp"] - sum(0.5 * (1 - v.real) for v in vec)) < 1e-12
    assert res["owned"] == [0, 2, 4, 6]
