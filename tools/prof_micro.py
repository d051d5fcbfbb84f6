#!/usr/bin/env python
"""tools/prof_micro.py [network] [reps] -- a network replayed as a compiled plan on resident inputs (ncu target for k_micro:
GHZ-1000 is 2 999 micro-steps in one launch)."""
import json
import os
import sys
ROOT = __file__.rsplit("/tools/", 1)[0]
sys.path.insert(0, ROOT)
os.environ["QTORCH_QUIET"] = "1"
import qtorch_b200 as qt
from qtorch_b200 import host_api

G = os.path.join(ROOT, "tests", "golden")
name = sys.argv[1] if len(sys.argv) > 1 else "ghz1000_zeros"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rec = json.load(open(os.path.join(G, "networks.json")))[name]
ranks, steps, inputs, flops = host_api.export_plan_linegraph(os.path.join(G, rec["qasm"]), os.path.join(G, rec["measure"]), os.path.join(G, rec["ordering"]), bool(rec["reduce"]))
eng = qt.Engine(0)
plan = eng.plan(ranks, steps)
plan.stage_inputs(0, inputs)
for _ in range(reps):
    eng.timer_start()
    plan.run_device_slot(0)
    ms = eng.timer_stop()
    print("%s: %d steps, %d launches, %.4f ms, value %r (golden %r)" % (name, len(steps), plan.launches, ms, complex(plan.read_output()[0]), rec["value"]))
