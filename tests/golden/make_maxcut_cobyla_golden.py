#!/usr/bin/env python
"""tests/golden/make_maxcut_cobyla_golden.py -- what the UNMODIFIED reference maxcutQAOA (oracle/_ref/maxcutQAOA_ref = src/maxcut.cpp +
vendored NLopt 2.4.2, LN_COBYLA, no stopping criterion) does on two small 3-regular graphs: how many objective evaluations it runs
until NLopt gives up ("roundoff-limited") and the angles it leaves in the angle file (the LAST evaluated ones, maxcut.cpp:199-202).
Writes tests/golden/maxcut_cobyla.json.  Run in the build container (the reference tree is needed to build the binary)."""
import json
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
EXE = os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "maxcutQAOA_ref")
out = {}
for graph, p in (("generated/prism6.dgf", 1), ("generated/cube8.dgf", 1), ("generated/prism6.dgf", 2)):
    with tempfile.TemporaryDirectory() as d:
        r = subprocess.run([EXE, os.path.join(HERE, graph), str(p), "0", "angles.txt"], cwd=d, capture_output=True, text=True, timeout=3600)
        n_edges = sum(1 for l in open(os.path.join(HERE, graph)) if l.startswith("e "))
        networks = r.stdout.count("Parsing nodes from file")
        angles = [float(x) for x in open(os.path.join(d, "angles.txt")).read().split()]
        msg = [l for l in r.stdout.splitlines() if l.startswith("nlopt")]
        key = "%s_p%d" % (os.path.basename(graph).split(".")[0], p)
        out[key] = {"graph": graph, "p": p, "evaluations": networks // n_edges, "last_angles": angles, "nlopt_message": msg[-1] if msg else ""}
        print(key, out[key], flush=True)
json.dump(out, open(os.path.join(HERE, "maxcut_cobyla.json"), "w"), indent=1)
