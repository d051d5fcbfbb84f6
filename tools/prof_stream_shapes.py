#!/usr/bin/env python
"""tools/prof_stream_shapes.py -- every bandwidth-bound step class of DESIGN.md section 3 on rank-13/14 tensors in ONE process
(one host array, CUDA-event timing, best of 4): kernel family, time, algorithmic GB/s and its fraction of the measured HBM peak."""
import json
import os
import sys
ROOT = __file__.rsplit("/tools/", 1)[0]
sys.path.insert(0, ROOT)
os.environ["QTORCH_QUIET"] = "1"
import numpy as np
import qtorch_b200 as qt

try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6546.9
KIND = {0: "k_micro", 1: "k_step_thread", 2: "k_gett", 3: "k_step_warp", 5: "k_dot/k_reduce", 6: "k_gett FUSE", 7: "k_apply"}
eng = qt.Engine(0)
rng = np.random.default_rng(0)
big = rng.standard_normal(4 ** 14) + 1j * rng.standard_normal(4 ** 14)
SHAPES = [
    (14, 14, list(range(14)), [9, 6, 8, 7, 0, 5, 13, 2, 10, 12, 4, 1, 3, 11]),   # closing inner product of config 2
    (14, 4, [3, 9], [0, 2]), (14, 2, [13], [0]), (14, 2, [0], [0]), (2, 14, [0], [7]), (2, 14, [1], [0]),
    (13, 1, [5], [0]), (13, 3, [2, 7], [0, 1]), (14, 2, [0, 1], [0, 1]),
    (13, 3, [12], [0]), (13, 3, [0], [2]), (3, 13, [1], [5]), (3, 13, [2], [0]), (4, 13, [0, 2], [3, 9]),
]
tensors = {}
def tensor_of(r, salt):
    key = (r, salt)
    if key not in tensors:
        tensors[key] = eng.tensor(r, big[salt:salt + 4 ** r] if r < 14 else big)
    return tensors[key]
for rA, rB, pA, pB in SHAPES:
    ta, tb = tensor_of(rA, 0), tensor_of(rB, 7)
    rC = rA + rB - 2 * len(pA)
    tc = eng.tensor(rC)
    eng.sync()
    eng.trace(True); eng.read_trace()
    best = 1e9
    for _ in range(4):
        eng.timer_start()
        eng.contract(ta, tb, pA, pB, out=tc)
        best = min(best, eng.timer_stop())
    tr = eng.read_trace(); eng.trace(False)
    kinds = sorted({KIND.get(r["kernel"], str(r["kernel"])) for r in tr})
    by = 16 * (4 ** rA + 4 ** rB + 4 ** rC)
    print("(%d,%d,k=%d -> %d) posA %s posB %s: %-14s %.3f ms  %.2f GB  %.0f GB/s  %.2f of %.0f" % (rA, rB, len(pA), rC, pA, pB, "+".join(kinds), best, by * 1e-9, by / best * 1e-6, by / best * 1e-6 / PEAK, PEAK), flush=True)
    tc.free()
    if rA == 14 and rB == 14:
        for t in tensors.values():
            t.free()
        tensors.clear()
