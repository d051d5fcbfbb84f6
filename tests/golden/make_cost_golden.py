#!/usr/bin/env python
"""tests/golden/make_cost_golden.py -> cost_plans.json

Golden plans of the reference's sampled greedy planner (ContractionTools::Contract(CostContractSimple, pValue),
/root/reference/src/ContractionTools.h:837-1046) for fixed seeds.  The reference seeds its generator from
std::random_device, so oracle/ref_harness's `cost` mode assigns the seed (see the access note in ref_harness.cpp);
everything else is the unmodified reference.  Needs /root/reference (build container only); the output is committed.
pValue = 1 only: for pValue >= 2 the reference spins forever once a super-node has absorbed all its neighbours (:1002)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

NETS = json.load(open(os.path.join(HERE, "networks.json")))
# (network record that supplies qasm + measurement, pValue, seed)
CASES = [("qft8_X8", 1, 7), ("qft8_X8", 1, 2024), ("testJW_YXXY", 1, 11), ("rand6_rxyz_stoch", 1, 3), ("qaoa20_node1_m125", 1, 5),
         ("arbitrary_gates", 1, 1), ("ghz64_zeros", 1, 9)]

out = {}
for name, p, seed in CASES:
    rec = NETS[name]
    cwd = os.path.join(HERE, rec.get("cwd", ""))
    qasm = os.path.join(HERE, rec["qasm"])
    if rec.get("cwd"):
        qasm = os.path.relpath(qasm, cwd)
    meas = os.path.join(HERE, rec["measure"])
    r = O.ref_harness(["cost", qasm, meas, p, seed, 8], cwd=cwd, timeout=300)
    assert "exception" not in r, (name, r.get("exception"))
    out["%s_p%d_s%d" % (name, p, seed)] = {
        "network": name, "p": p, "seed": seed, "value": [float(r["value"][0]), float(r["value"][1])],
        "flops": int(r["flops"][0]), "nodes": int(r["nodes"][0]), "plan": r["plan"]}
    print(name, p, seed, r["value"], r["flops"][0], len(r["plan"]))
json.dump(out, open(os.path.join(HERE, "cost_plans.json"), "w"), indent=1, sort_keys=True)
