#!/usr/bin/env python
"""tools/check_tma.py -- the TMA-fed tile kernel against the oracle on the big-step shapes, and its time next to the cp.async gather"""
import os
import sys
os.environ.setdefault("QTB_TMA", "1")
ROOT = __file__.rsplit("/tools/", 1)[0]
sys.path.insert(0, ROOT)
import numpy as np
import qtorch_b200 as qt
from oracle import oracle as O

eng = qt.Engine(0)
rng = np.random.default_rng(1)
O.lib().qto_set_threads(16)
# small versions of the config-2 / config-4 leg maps (ranks cut down so the oracle finishes), then the real ones timed
CASES = [(8, 8, [0, 2, 5], [1, 7, 6]), (7, 9, [0, 4, 6], [6, 8, 5]), (6, 10, [0, 2, 3], [2, 7, 9]), (6, 10, [0, 1, 4], [9, 6, 7]),
         (7, 7, [0, 3], [5, 1]), (8, 6, [0, 4], [2, 5]), (9, 7, [1, 4, 5], [5, 2, 3]), (8, 8, [1, 3, 4, 6], [2, 7, 4, 1])]
for rA, rB, pA, pB in CASES:
    A = rng.standard_normal(4 ** rA) + 1j * rng.standard_normal(4 ** rA)
    B = rng.standard_normal(4 ** rB) + 1j * rng.standard_normal(4 ** rB)
    ta, tb = eng.tensor(rA, A), eng.tensor(rB, B)
    eng.reset_stats()
    C = eng.contract(ta, tb, pA, pB).download()
    st = eng.stats()
    ref = O.contract(A, rA, B, rB, pA, pB)
    err = np.abs(C - ref).max() / max(1.0, np.abs(ref).max())
    print("(%d,%d,k=%d) posA %s posB %s: tma_launches %d, err %.2e %s" % (rA, rB, len(pA), pA, pB, st["tma_launches"], err, "OK" if err < 1e-12 else "WRONG"), flush=True)
    ta.free(); tb.free()
