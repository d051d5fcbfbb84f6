#!/usr/bin/env python
"""tools/prof_maxcut_ncu.py -- three evaluations of the p=2 objective of config 3 without CUDA graphs (ncu target for the batch kernel
k_micro_t<512> on thread-block clusters: `QTB_PLAN_GRAPH=0 ncu --set full -k regex:k_micro_t -s 2 -c 1 python tools/prof_maxcut_ncu.py`)."""
import json
import os
import sys
ROOT = __file__.rsplit("/tools/", 1)[0]
sys.path.insert(0, ROOT)
os.environ["QTORCH_QUIET"] = "1"
os.environ.setdefault("QTB_PLAN_GRAPH", "0")
from qtorch_b200 import host_api
G = os.path.join(ROOT, "tests", "golden")
rec = json.load(open(os.path.join(G, "maxcut.json")))["3reg30_p2_default"]
q = host_api.QaoaObjective(os.path.join(G, rec["graph"]), 2)
for _ in range(3):
    vals, fp = q.evaluate(rec["betas_gammas"])
print("fp", fp, "golden", rec["fp"])
q.close()
