#!/usr/bin/env python
"""tools/prof_maxcut.py -- config 3: objective evaluations per second of the QAOA term batch (qtb_batch_*), p = 1 and 2,
CUDA-event timed, plus the values against tests/golden/maxcut.json."""
import json
import os
import sys
import time
ROOT = __file__.rsplit("/tools/", 1)[0]
sys.path.insert(0, ROOT)
os.environ["QTORCH_QUIET"] = "1"
import numpy as np
from qtorch_b200 import host_api

G = os.path.join(ROOT, "tests", "golden")
MC = json.load(open(os.path.join(G, "maxcut.json")))
eng = host_api.engine()
for name in ("3reg30_p1_default", "3reg30_p2_default"):
    rec = MC[name]
    t0 = time.perf_counter()
    q = host_api.QaoaObjective(os.path.join(G, rec["graph"]), rec["p"])
    build = time.perf_counter() - t0
    vals, fp = q.evaluate(rec["betas_gammas"])
    ref = np.array([complex(*t) for t in rec["terms"]])
    err = np.abs(vals - ref).max()
    for _ in range(20):
        q.evaluate(rec["betas_gammas"])
    K = 300
    eng.sync(); t0 = time.perf_counter(); eng.timer_start()
    for i in range(K):
        q.evaluate([a * (1.0 + 1e-3 * (i % 7)) for a in rec["betas_gammas"]])
    ms = eng.timer_stop() / K
    wall = (time.perf_counter() - t0) * 1e3 / K
    print("%s: %d terms, units/eval %d, launches/eval %d, plan build %.2f s, max|err| %.2e, fp err %.2e: %.4f ms/eval (events), %.4f ms wall -> %.0f terms/s"
          % (name, len(vals), q.units, q.launches, build, err, abs(fp - rec["fp"]), ms, wall, len(vals) / (ms * 1e-3)), flush=True)
    q.close()

# what bounds an evaluation: every edge's light-cone network runs as ONE CTA of the graph's single launch, so an evaluation lasts
# as long as its slowest term -- evaluate each term ALONE (a "rank" that owns one edge) and compare with all 45 together
for name in (("3reg30_p1_default", "3reg30_p2_default") if "--per-term" in sys.argv else ()):
    rec = MC[name]
    n = len(rec["terms"])
    alone = []
    for r in range(n):
        q = host_api.QaoaObjective(os.path.join(G, rec["graph"]), rec["p"], rank=r, world=n)
        for _ in range(5):
            q.evaluate(rec["betas_gammas"])
        eng.timer_start()
        for _ in range(50):
            q.evaluate(rec["betas_gammas"])
        alone.append((eng.timer_stop() / 50, q.units))
        q.close()
    ms = sorted(a[0] for a in alone)
    print("%s: one term alone: min %.4f  median %.4f  max %.4f ms (units of the slowest term %d); sum over terms %.3f ms"
          % (name, ms[0], ms[len(ms) // 2], ms[-1], max(alone)[1], sum(ms)), flush=True)
