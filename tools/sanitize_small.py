#!/usr/bin/env python
"""tools/sanitize_small.py -- a few small invocations of every kernel family for `compute-sanitizer --tool memcheck`
(tile kernel plain + fused, the opt-in shared-sum kernel when QTB_GETT_C1=5, streaming kernel, split-K, grouped micro-steps on one CTA
and -- QTB_MICRO_CLUSTER=4 -- on a cluster, a term batch)."""
import json
import os
import sys
ROOT = __file__.rsplit("/tools/", 1)[0]
sys.path.insert(0, ROOT)
os.environ["QTORCH_QUIET"] = "1"
import numpy as np
import qtorch_b200 as qt
from qtorch_b200 import host_api
from oracle import oracle as O

eng = host_api.engine()
e2 = qt.Engine(0)
rng = np.random.default_rng(3)
O.lib().qto_set_threads(8)
bad = 0
for rA, rB, pA, pB in [(7, 7, [0, 3], [5, 1]), (7, 7, [1, 4, 6], [0, 3, 2]), (8, 2, [3], [0]), (8, 1, [0], [0]), (6, 6, [0, 1, 2, 3, 4, 5], [5, 3, 1, 0, 2, 4]), (4, 3, [1], [2])]:
    A = rng.standard_normal(4 ** rA) + 1j * rng.standard_normal(4 ** rA)
    B = rng.standard_normal(4 ** rB) + 1j * rng.standard_normal(4 ** rB)
    C = e2.contract(e2.tensor(rA, A), e2.tensor(rB, B), pA, pB).download()
    ref = O.contract(A, rA, B, rB, pA, pB)
    err = np.abs(C - ref).max() / max(1.0, np.abs(ref).max())
    bad += err > 1e-12
    print("step (%d,%d,k=%d): err %.1e" % (rA, rB, len(pA), err), flush=True)
# fused tile kernel + inner product (rank-10 intermediate)
rA, rB, pA, pB = 7, 7, [0, 3], [5, 1]
A = rng.standard_normal(4 ** rA) + 1j * rng.standard_normal(4 ** rA)
B = rng.standard_normal(4 ** rB) + 1j * rng.standard_normal(4 ** rB)
D = rng.standard_normal(4 ** 10) + 1j * rng.standard_normal(4 ** 10)
perm = np.random.default_rng(5).permutation(10).tolist()
T = O.contract(A, rA, B, rB, pA, pB)
ref = O.contract(T, 10, D, 10, list(range(10)), perm)[0]
tt = e2.contract(e2.tensor(rA, A), e2.tensor(rB, B), pA, pB)
val = e2.contract(tt, e2.tensor(10, D), list(range(10)), perm).scalar()
bad += abs(val - ref) > 1e-11 * max(1.0, abs(ref))
print("fused: err %.1e" % abs(val - ref), flush=True)
# grouped micro-steps through the host mirror + a term batch
G = os.path.join(ROOT, "tests", "golden")
nets = json.load(open(os.path.join(G, "networks.json")))
for name in ("qft8_X8", "ghz64_zeros"):
    rec = nets[name]
    v, _, _, _ = host_api.contract_linegraph(*[os.path.join(G, rec[k]) for k in ("qasm", "measure", "ordering")], True)
    bad += abs(v - complex(*rec["value"])) > 1e-10
    print(name, "err %.1e" % abs(v - complex(*rec["value"])), flush=True)
mc = json.load(open(os.path.join(G, "maxcut.json")))["3reg30_p2_default"]
q = host_api.QaoaObjective(os.path.join(G, mc["graph"]), 2, rank=0, world=9)
vals, _ = q.evaluate(mc["betas_gammas"])
err = max(abs(v - complex(*mc["terms"][e])) for e, v in zip(q.owned, vals))
bad += err > 1e-10
print("p=2 term batch (5 terms): err %.1e" % err, flush=True)
q.close()
print("sanitize_small failures:", bad)
sys.exit(1 if bad else 0)
