#!/usr/bin/env python
"""tools/micro_timeline.py [term ...] -- debugging aid: per-level clock stamps of the grouped launch (k_micro) of single QAOA p=2 terms
(QTB_MICRO_TIMELINE=1 makes compiled plans carry a stamp array; qtb_debug_dump_micro_timelines prints them)."""
import ctypes
import json
import os
import sys
ROOT = __file__.rsplit("/tools/", 1)[0]
sys.path.insert(0, ROOT)
os.environ["QTORCH_QUIET"] = "1"
os.environ["QTB_MICRO_TIMELINE"] = "1"
import qtorch_b200 as qt
from qtorch_b200 import host_api

G = os.path.join(ROOT, "tests", "golden")
rec = json.load(open(os.path.join(G, "maxcut.json")))["3reg30_p2_default"]
terms = [int(x) for x in sys.argv[1:]] or [18]
for r in terms:
    q = host_api.QaoaObjective(os.path.join(G, rec["graph"]), 2, rank=r, world=45)
    for _ in range(5):
        q.evaluate(rec["betas_gammas"])
    print("term", r, "units", q.units, file=sys.stderr)
    q.evaluate(rec["betas_gammas"])
    qt.load_library().qtb_debug_dump_micro_timelines(ctypes.c_int(100))
    q.close()
