#!/usr/bin/env python
"""tools/check_g3.py [--time-only | --parity-only] -- the compute-bound tile kernel selected by QTB_GETT_C1 (5 = the shared-sum
3M kernel of gett3m.cuh, 2 = k_gett 3M, the default) against the oracle (plain and fused with the inner product
that follows), then its time on the rank-14 steps of BASELINE config 2.  The variant comes from the environment
(QTB_GETT_C1 / QTB_G3, read once per process)."""
import os
import sys
ROOT = __file__.rsplit("/tools/", 1)[0]
sys.path.insert(0, ROOT)
import numpy as np
import qtorch_b200 as qt
from oracle import oracle as O

quick = "--time-only" in sys.argv
eng = qt.Engine(0)
rng = np.random.default_rng(1)
O.lib().qto_set_threads(16)
label = "C1=%s G3=%s" % (os.environ.get("QTB_GETT_C1", "2"), os.environ.get("QTB_G3", "0"))
CASES = [(8, 8, [0, 2, 5], [1, 7, 6]), (7, 9, [0, 4, 6], [6, 8, 5]), (6, 10, [0, 2, 3], [2, 7, 9]), (6, 10, [0, 1, 4], [9, 6, 7]),
         (7, 7, [0, 3], [5, 1]), (8, 6, [0, 4], [2, 5]), (9, 7, [1, 4, 5], [5, 2, 3]), (8, 8, [1, 3, 4, 6], [2, 7, 4, 1]),
         (10, 6, [2, 5, 9], [0, 2, 3]), (7, 8, [0, 1], [0, 3]), (7, 7, [5, 6], [0, 1]), (8, 7, [0, 1], [5, 6]), (7, 8, [2, 6], [0, 1])]
bad = 0
if not quick:
    for rA, rB, pA, pB in CASES:
        A = rng.standard_normal(4 ** rA) + 1j * rng.standard_normal(4 ** rA)
        B = rng.standard_normal(4 ** rB) + 1j * rng.standard_normal(4 ** rB)
        ta, tb = eng.tensor(rA, A), eng.tensor(rB, B)
        C = eng.contract(ta, tb, pA, pB).download()
        ref = O.contract(A, rA, B, rB, pA, pB)
        err = np.abs(C - ref).max() / max(1.0, np.abs(ref).max())
        bad += err >= 1e-12
        print("[%s] (%d,%d,k=%d) posA %s posB %s: err %.2e %s" % (label, rA, rB, len(pA), pA, pB, err, "OK" if err < 1e-12 else "WRONG"), flush=True)
        ta.free(); tb.free()
    # fused with the inner product that follows (rank-10 intermediate), both operand orders
    for (rA, rB, pA, pB) in [(7, 7, [0, 3], [5, 1]), (6, 8, [1, 2], [7, 0]), (8, 6, [0, 4], [2, 5]), (7, 9, [0, 4, 6], [6, 8, 5])]:
        for t_is_a in (True, False):
            A = rng.standard_normal(4 ** rA) + 1j * rng.standard_normal(4 ** rA)
            B = rng.standard_normal(4 ** rB) + 1j * rng.standard_normal(4 ** rB)
            rT = rA + rB - 2 * len(pA)
            D = rng.standard_normal(4 ** rT) + 1j * rng.standard_normal(4 ** rT)
            perm = np.random.default_rng(5).permutation(rT).tolist()
            T = O.contract(A, rA, B, rB, pA, pB)
            if t_is_a:
                posA2, posB2 = list(range(rT)), perm
                ref = O.contract(T, rT, D, rT, posA2, posB2)
            else:
                posA2, posB2 = list(range(rT)), np.argsort(perm).tolist()
                ref = O.contract(D, rT, T, rT, posA2, posB2)
            ta, tb, td = eng.tensor(rA, A), eng.tensor(rB, B), eng.tensor(rT, D)
            before = eng.stats()["launches"]
            tt = eng.contract(ta, tb, pA, pB)
            out = eng.contract(tt, td, posA2, posB2) if t_is_a else eng.contract(td, tt, posA2, posB2)
            val = out.scalar()
            n = eng.stats()["launches"] - before
            err = abs(val - ref[0]) / max(1.0, abs(ref[0]))
            bad += err >= 1e-11
            print("[%s] fused (%d,%d,k=%d) tIsA=%s: launches %d err %.2e %s" % (label, rA, rB, len(pA), t_is_a, n, err, "OK" if err < 1e-11 else "WRONG"), flush=True)
            for t in (ta, tb, td, tt, out):
                t.free()
print("[%s] parity failures: %d" % (label, bad), flush=True)
if "--parity-only" in sys.argv:
    sys.exit(1 if bad else 0)

# timing: the three rank-14 steps of config 2 and the fused closing pair
for rA, rB, pA, pB in [(10, 10, [0, 2, 5], [1, 7, 6]), (9, 11, [0, 4, 6], [6, 8, 5]), (6, 14, [0, 2, 3], [2, 7, 9])]:
    ta = eng.tensor(rA, rng.standard_normal(4 ** rA) + 1j * rng.standard_normal(4 ** rA))
    tb = eng.tensor(rB, rng.standard_normal(4 ** rB) + 1j * rng.standard_normal(4 ** rB))
    tc = eng.tensor(rA + rB - 2 * len(pA))
    eng.sync()
    best = 1e9
    for _ in range(4):
        eng.timer_start()
        eng.contract(ta, tb, pA, pB, out=tc)
        best = min(best, eng.timer_stop())
    print("[%s] step (%d,%d,k=3 -> 14): %.3f ms  %.2f TFLOP/s" % (label, rA, rB, best, 8 * 4 ** 17 / best * 1e-9), flush=True)
    ta.free(); tb.free(); tc.free()
rA, rB, pA, pB = 6, 14, [1, 3, 5], [0, 6, 12]
ta = eng.tensor(rA, rng.standard_normal(4 ** rA) + 1j * rng.standard_normal(4 ** rA))
big = rng.standard_normal(4 ** rB) + 1j * rng.standard_normal(4 ** rB)
tb = eng.tensor(rB, big)
td = eng.tensor(14, big[::-1].copy())
perm = np.random.default_rng(5).permutation(14).tolist()
eng.sync()
best = 1e9
for _ in range(4):
    eng.timer_start()
    tt = eng.contract(ta, tb, pA, pB)
    out = eng.contract(tt, td, list(range(14)), perm)
    val = out.scalar()
    best = min(best, eng.timer_stop())
    tt.free(); out.free()
print("[%s] fused (6,14,k=3 -> 14).(14,14 -> 0): %.3f ms" % (label, best), flush=True)
