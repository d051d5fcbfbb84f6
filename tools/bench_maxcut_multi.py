#!/usr/bin/env python
"""tools/bench_maxcut_multi.py [--p P] [--evals K] -- BASELINE config 3 (maxcutQAOA objective on 3regRand30Node50.dgf) with
the edges dealt round-robin over the ranks of a torchrun job: every rank evaluates its <ZiZj> terms in grouped launches,
one NCCL allreduce of the partial objective per evaluation (qtb_allreduce_sum).  One JSON line from rank 0."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["QTORCH_QUIET"] = "1"

ap = argparse.ArgumentParser()
ap.add_argument("--p", type=int, default=1)
ap.add_argument("--evals", type=int, default=100)
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
os.environ.setdefault("QTORCH_DEVICE", str(local))

import numpy as np  # noqa: E402
import torch  # noqa: E402
from qtorch_b200 import host_api  # noqa: E402

torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
G = os.path.join(ROOT, "tests", "golden")
rec = json.load(open(os.path.join(G, "maxcut.json")))["3reg30_p%d_default" % args.p]
eng = host_api.engine()
q = host_api.QaoaObjective(os.path.join(G, rec["graph"]), args.p, rank=rank, world=world)
if dist is not None:
    uid = [eng.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(world, rank, uid[0])


def objective(bg):
    vals, part = q.evaluate(bg)
    if dist is not None:
        part = eng.allreduce_sum(np.array([part], dtype=np.complex128))[0].real
    return part


bg = list(rec["betas_gammas"])
fp = objective(bg)
ok = abs(fp - rec["fp"]) <= 1e-10 * max(1.0, abs(rec["fp"]))
for _ in range(5):
    objective([a * 0.9 for a in bg])
if dist is not None:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(args.evals):
    objective([a * (1.0 + 1e-3 * i) for a in bg])            # new angles every evaluation, as an optimiser would ask
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if dist is not None:
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t[0])
if rank == 0:
    n_edges = len(rec["terms"])
    print(json.dumps({"config": "maxcutQAOA 3regRand30Node50 p=%d" % args.p, "n_gpus": world, "edges": n_edges, "evaluations": args.evals,
                      "ms_per_evaluation": dt / args.evals * 1e3, "terms_per_s": n_edges * args.evals / dt,
                      "objective_matches_reference_1e-10": bool(ok), "launches_per_evaluation_rank0": q.launches,
                      "timing": "host wall clock around K objective evaluations incl. allreduce, max over ranks"}))
if dist is not None:
    dist.destroy_process_group()
