#!/usr/bin/env python
"""tools/prof_fused.py [reps] -- the closing pair of BASELINE config 2 as the engine runs it: (6,14,k=3 -> 14) fused with the
(14,14,k=14 -> 0) inner product that follows (one k_gett FUSE launch + the partial-sum reduction).  ncu target."""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import qtorch_b200 as qt

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
eng = qt.Engine(0)
rng = np.random.default_rng(0)
rA, rB, pA, pB = 6, 14, [1, 3, 5], [0, 6, 12]
ta = eng.tensor(rA, rng.standard_normal(4 ** rA) + 1j * rng.standard_normal(4 ** rA))
big = rng.standard_normal(4 ** rB) + 1j * rng.standard_normal(4 ** rB)
tb = eng.tensor(rB, big)
td = eng.tensor(14, big[::-1].copy())
perm = np.random.default_rng(5).permutation(14).tolist()
eng.sync()
for _ in range(reps):
    eng.timer_start()
    tt = eng.contract(ta, tb, pA, pB)
    out = eng.contract(tt, td, list(range(14)), perm)
    val = out.scalar()
    ms = eng.timer_stop()
    print("fused (6,14,k=3 -> 14) . (14,14,k=14 -> 0): %.3f ms  value %r" % (ms, val))
    tt.free(); out.free()
