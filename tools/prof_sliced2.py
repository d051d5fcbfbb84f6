#!/usr/bin/env python
"""tools/prof_sliced2.py -- the sliced-amplitude executor (qtb_sliced_*) on the config-2 term and the config-4 amplitude:
ms per amplitude for 1/2/3 lanes, one amplitude at a time vs two in flight, and the share a rank of an 8-GPU job would own."""
import json
import os
import sys
import time
ROOT = __file__.rsplit("/tools/", 1)[0]
sys.path.insert(0, ROOT)
os.environ["QTORCH_QUIET"] = "1"
from qtorch_b200 import host_api

G = os.path.join(ROOT, "tests", "golden")
nets = json.load(open(os.path.join(G, "networks.json")))
eng = host_api.engine()
which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["qaoa30_z27z29", "rand42_cn4_d20_zeros"]
SHARES = [(1, 0, 2), (1, 0, 4), (2, 0, 8), (2, 3, 8)] if "--shares" in sys.argv else [(0, 0, 1), (1, 0, 1), (2, 0, 1), (1, 0, 2), (1, 0, 4), (2, 0, 8), (2, 3, 8)]
for name in which:
    rec = nets[name]
    paths = [os.path.join(G, rec[k]) for k in ("qasm", "measure", "ordering")]
    ref = complex(*rec["value"])
    for s, rank, world in SHARES:
        for lanes in (1, 2, 3):
            net = host_api.SlicedNetwork(*paths, True, slice_wires=s, lanes=lanes, rank=rank, world=1 if world == 1 else -world)
            # world > 1 here only selects this rank's share of the slices (no communicator: the reduction is skipped below)
            net.stage(0); net.stage(1)
            try:
                v = net.end(net.begin(0))
            except Exception as e:
                print(name, s, lanes, "failed", e); net.close(); continue
            K = 10
            eng.sync(); t0 = time.perf_counter(); eng.timer_start()
            for _ in range(K):
                v = net.end(net.begin(0))
            ms_seq = eng.timer_stop() / K
            wall_seq = (time.perf_counter() - t0) * 1e3 / K
            eng.timer_start()
            t = [net.begin(0), net.begin(1)]
            for i in range(K):
                v = net.end(t[i % 2])
                t[i % 2] = net.begin(i % 2)
            net.end(t[0]); net.end(t[1])
            ms_pipe = eng.timer_stop() / (K + 2)
            ok = (world > 1) or abs(v - ref) <= 1e-10 * max(1, abs(ref))
            print("%s s=%d share=%d/%d lanes=%d owned=%d inv=%d/%d peak=%d launches/slice=%d(prefix %d): %.3f ms/amp sequential (wall %.3f), %.3f ms/amp two in flight  ok=%s"
                  % (name, s, rank, world, lanes, net.owned, net.invariant_steps, net.steps, net.peak_rank, net.launches_per_slice, net.launches_prefix,
                     ms_seq, wall_seq, ms_pipe, ok), flush=True)
            net.close()
            if net.owned <= 1 and lanes == 1 and s > 0:
                pass
