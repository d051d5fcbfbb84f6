#!/usr/bin/env python
"""tools/prof_sliced.py [slice_wires] [network] -- where one index-sliced amplitude (default: the config-2 term) spends its time: per-segment
CUDA-event trace of the invariant prefix (run once) and of one pass over the per-slice suffix, then the untraced timing."""
import json
import os
import sys
ROOT = __file__.rsplit("/tools/", 1)[0]
sys.path.insert(0, ROOT)
import qtorch_b200 as qt
from qtorch_b200 import host_api, slicing

G = os.path.join(ROOT, "tests", "golden")
rec = json.load(open(os.path.join(G, "networks.json")))[sys.argv[2] if len(sys.argv) > 2 else "qaoa30_z27z29"]
s = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ranks, steps, inputs, flops = host_api.export_plan_linegraph(os.path.join(G, rec["qasm"]), os.path.join(G, rec["measure"]), os.path.join(G, rec["ordering"]), True)
wires = slicing.choose_wires(ranks, steps, s)
eng = qt.Engine(0)
plan, cuts, n_inv = slicing.compile_sliced(eng, ranks, steps, wires)
sl = slicing.all_slices(wires)
for slot, d in enumerate(sl):
    plan.stage_inputs(slot, slicing.slice_inputs(inputs, ranks, cuts, wires, d))
print("slices %d, invariant steps %d of %d, units shared %.3e, per slice %.3e, launches/plan %d" % (len(sl), n_inv, len(steps), plan.prefix_units, plan.units - plan.prefix_units, plan.launches))
val = plan.run_slots(range(len(sl)))
print("value", val, "golden", rec["value"], "ok", abs(val - complex(*rec["value"])) <= 1e-10)
eng.trace(True)
plan.run_slots([0])
tr = eng.read_trace()
eng.trace(False)
names = {0: "micro", 1: "thread", 2: "gett", 3: "warp", 5: "reduce", 6: "fused", 7: "apply"}
tot = 0.0
for t in tr:
    tot += t["ms"]
    print("  %-6s rA=%2d rB=%2d k=%3d  %.4f ms" % (names.get(t["kernel"], t["kernel"]), t["rank_a"], t["rank_b"], t["k"], t["ms"]))
print("traced total %.3f ms (prefix + one slice)" % tot)
for reps in (1, 3):
    eng.timer_start()
    for _ in range(reps):
        plan.run_slots(range(len(sl)))
    print("run_slots x%d over %d slices: %.3f ms per amplitude" % (reps, len(sl), eng.timer_stop() / reps))
eng.timer_start()
plan.run_slots([0, 1])
print("2 slices (8-GPU share): %.3f ms" % eng.timer_stop())
plan.destroy()
