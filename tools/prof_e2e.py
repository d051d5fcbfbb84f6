#!/usr/bin/env python
"""tools/prof_e2e.py -- where the end-to-end (files -> host mirror -> device -> scalar) call of config 2 spends its time:
per-launch trace of one eager term, blocking vs pipelined timing, host-only (plan-only) time of the bookkeeping."""
import json
import os
import sys
import time
ROOT = __file__.rsplit("/tools/", 1)[0]
sys.path.insert(0, ROOT)
os.environ["QTORCH_QUIET"] = "1"
import qtorch_b200 as qt
from qtorch_b200 import host_api

G = os.path.join(ROOT, "tests", "golden")
rec = json.load(open(os.path.join(G, "networks.json")))["qaoa30_z27z29"]
args = (os.path.join(G, rec["qasm"]), os.path.join(G, rec["measure"]), os.path.join(G, rec["ordering"]), True)
eng = host_api.engine()
for _ in range(3):
    v = host_api.contract_linegraph(*args)[0]
print("value", v)
eng.trace(True); eng.read_trace()
host_api.contract_linegraph(*args)
tr = eng.read_trace(); eng.trace(False)
names = {0: "micro", 1: "thread", 2: "gett", 3: "warp", 5: "reduce", 6: "fused"}
by = {}
for t in tr:
    k = names.get(t["kernel"], t["kernel"]); by.setdefault(k, [0, 0.0]); by[k][0] += 1; by[k][1] += t["ms"]
print("eager term trace:", {k: (n, round(ms, 3)) for k, (n, ms) in by.items()}, "sum %.3f ms" % sum(ms for _, ms in by.values()))
for t in tr:
    if t["ms"] > 0.05: print("   %-6s rA=%2d rB=%2d k=%3d %.3f ms" % (names.get(t["kernel"]), t["rank_a"], t["rank_b"], t["k"], t["ms"]))
K = 10
eng.timer_start(); t0 = time.time()
for _ in range(K): host_api.contract_linegraph(*args)
print("blocking : %.3f ms/term (device events)  %.3f ms wall" % (eng.timer_stop() / K, (time.time() - t0) * 1e3 / K))
eng.timer_start(); t0 = time.time()
job = host_api.LinegraphJob(*args)
tb = 0.0
for i in range(K):
    t1 = time.time()
    nxt = host_api.LinegraphJob(*args) if i + 1 < K else None
    tb += time.time() - t1
    job.result(); job = nxt
print("pipelined: %.3f ms/term (device events)  %.3f ms wall; begin() host time %.3f ms" % (eng.timer_stop() / K, (time.time() - t0) * 1e3 / K, tb * 1e3 / (K - 1)))
t0 = time.time()
for _ in range(K): host_api.export_plan_linegraph(*args)
print("plan-only host bookkeeping incl. export: %.3f ms" % ((time.time() - t0) * 1e3 / K))
print(eng.stats())
