#!/usr/bin/env python
"""tools/prof_step.py rA rB k posA.. posB.. [reps] -- run one contraction step repeatedly (ncu target)."""
import sys
import numpy as np
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import qtorch_b200 as qt

a = [int(x) for x in sys.argv[1:]]
rA, rB, k = a[:3]
pA, pB = a[3:3 + k], a[3 + k:3 + 2 * k]
reps = a[3 + 2 * k] if len(a) > 3 + 2 * k else 3
eng = qt.Engine(0)
rng = np.random.default_rng(0)
ta = eng.tensor(rA, rng.standard_normal(4 ** rA) + 1j * rng.standard_normal(4 ** rA))
tb = eng.tensor(rB, rng.standard_normal(4 ** rB) + 1j * rng.standard_normal(4 ** rB))
tc = eng.tensor(rA + rB - 2 * k)
eng.sync()
for _ in range(reps):
    eng.timer_start()
    eng.contract(ta, tb, pA, pB, out=tc)
    ms = eng.timer_stop()
    U = 4 ** (rA + rB - k)
    by = 16 * (4 ** rA + 4 ** rB + 4 ** (rA + rB - 2 * k))
    print("step (%d,%d,k=%d): %.3f ms  %.2f TFLOP/s  %.1f GB/s" % (rA, rB, k, ms, 8 * U / ms * 1e-9, by / ms * 1e-6))
