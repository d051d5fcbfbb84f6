#!/usr/bin/env python
"""bench.py -- QAOA <Z_i Z_j> terms/s on B200 (BASELINE.json metric), workload = BASELINE config 2.

Workload (named in config.workload): p=1 QAOA MaxCut expectation terms of Samples/4regRand30Node5-p1.qasm
(30 qubits, 4-regular, 60 edges), line-graph ordering frozen in tests/golden/orderings/qaoa30_z27z29.qbb.out
(one ordering serves every term: the measurement caps do not change the line graph).  One "step" = one full
term contraction per rank: 299 pairwise steps, 6.935e10 units (5.5e11 flop), four rank-14 DMMA steps + one
268M-term inner product.  Terms are dealt round-robin to ranks (term = step*N + rank mod 60), no data-path
collective; the per-step scalars are summed with one NCCL allreduce at the end ("scaling": "weak").

  value : device-resident: the compiled plan (qtb_plan_*) with every term's inputs staged in HBM beforehand
  e2e   : the user's call through the C++ host mirror (Network -> ReduceCircuit -> LGContract -> GetFinalValue)
          with .qasm / measurement files in, gate tensors uploaded (H2D) and the scalar read back (D2H) per term

--impl reference : the UNMODIFIED reference (oracle/_ref/ref_harness, built from /root/reference/src) replaying the
same plan on the host cores, each step a bounded sample of the term (plan steps 0..294, all but the four rank-14
steps and the final inner product), extrapolated by units to terms/s.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
QASM = os.path.join(GOLDEN, "Samples", "4regRand30Node5-p1.qasm")
ORDERING = os.path.join(GOLDEN, "orderings", "qaoa30_z27z29.qbb.out")
NETS = os.path.join(GOLDEN, "networks.json")
N_QUBITS = 30
UNITS_PER_TERM = 69351174176          # getNumFloatOps() of one term (reference Network.h:884-885), golden
FP64_PEAK_TFLOPS = 37.1               # measured on this pool's B200: DMMA m8n8k4 probe, profiles/r01_probe_fp64.jsonl
METRIC = "qaoa_zz_terms_per_s"
WORKLOAD = "cfg2: 4regRand30Node5-p1.qasm (30q, 4-regular) <ZiZj> terms, frozen linegraph-qbb plan, 299 steps, 6.935e10 units/term"


def edges_of_circuit():
    """the 60 graph edges: distinct CNOT pairs of the p=1 circuit, in file order"""
    seen, out = set(), []
    for line in open(QASM):
        t = line.split()
        if len(t) == 3 and t[0] == "CNOT":
            e = (int(t[1]), int(t[2]))
            if e not in seen and (e[1], e[0]) not in seen:
                seen.add(e)
                out.append(e)
    return out


def write_measure_file(directory, edge):
    m = ["T"] * N_QUBITS
    m[edge[0]] = "Z"
    m[edge[1]] = "Z"
    path = os.path.join(directory, "zz_%d_%d.txt" % edge)
    with open(path, "w") as f:
        f.write(" ".join(m) + "\n")
    return path


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe): one streaming
    nvidia-smi process (-lms 100) started before and killed after."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        fd, self.path = tempfile.mkstemp(prefix="qtb_clocks_", suffix=".csv")
        os.close(fd)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        samples = []
        if self.path and os.path.exists(self.path):
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) >= 7:
                    samples.append(f)
        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = sorted(v for v in (num(s[0]) for s in samples) if v is not None)
        mx = [v for v in (num(s[1]) for s in samples) if v is not None]
        pw = [v for v in (num(s[2]) for s in samples) if v is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "reasons": reasons, "samples": len(samples)}


def cpu_reference_sample(threads, budget_units=3.0e8):
    """Time the unmodified reference on the box's host cores over a bounded sample of the term (see module doc).
    Returns dict(value=terms/s extrapolated by units, seconds, units, steps)."""
    from oracle import oracle as O
    if not O.ref_available():
        return None
    nets = json.load(open(NETS))
    rec = nets["qaoa30_z27z29"]
    with tempfile.TemporaryDirectory() as d:
        plan = os.path.join(d, "plan.txt")
        with open(plan, "w") as f:
            for p in rec["plan"]:
                f.write("%s %s\n" % tuple(p.split(",")))
        meas = os.path.join(GOLDEN, rec["measure"])
        out = O.ref_harness(["seq", QASM, meas, plan, threads, budget_units], cwd=GOLDEN, timeout=900)
    units, secs, steps = int(out["flops"][0]), float(out["seconds"][0]), int(out["steps"][0])
    return {"value": (units / secs) / UNITS_PER_TERM, "seconds": secs, "units": units, "steps": steps}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(threads)
    times, last = [], None
    for _ in range(args.steps):
        last = cpu_reference_sample(threads)
        if last is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_harness not built (needs /root/reference at build time)"}))
            return
        times.append(last["seconds"] * UNITS_PER_TERM / last["units"])       # extrapolated seconds per term
    sec_per_term = sum(times) / len(times)
    value = 1.0 / sec_per_term
    sample = "plan steps 0..%d of 299 (%.3g of %.3g units) per step, extrapolated by units" % (last["steps"] - 1, last["units"], UNITS_PER_TERM)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "terms/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec_per_term * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "reference_build": "g++ -O2 -std=c++11 -pthread, unmodified /root/reference/src via oracle/ref_harness.cpp"},
        "cpu_baseline": {"value": value, "unit": "terms/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "terms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("QTORCH_DEVICE", str(local))
    os.environ["QTORCH_QUIET"] = "1"

    import numpy as np
    import torch
    import qtorch_b200 as qt
    from qtorch_b200 import host_api

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    eng = host_api.engine()                 # the host mirror's engine context (shared by both measured paths)
    edges = edges_of_circuit()
    assert len(edges) == 60
    tmp = tempfile.mkdtemp(prefix="qtb_bench_")
    meas_files = [write_measure_file(tmp, e) for e in edges]
    W, K = args.warmup, args.steps
    my_terms = [(s * world + rank) % len(edges) for s in range(W + K)]

    # ---------------- device-resident path: one compiled plan, per-term input sets staged in HBM ----------------
    ranks, steps, inputs0, flops = host_api.export_plan_linegraph(QASM, meas_files[my_terms[0]], ORDERING, True)
    assert flops == UNITS_PER_TERM
    plan = eng.plan(ranks, steps)
    slots = {}
    for t in sorted(set(my_terms)):
        _, _, inp, _ = host_api.export_plan_linegraph(QASM, meas_files[t], ORDERING, True)
        slots[t] = len(slots)
        plan.stage_inputs(slots[t], inp)
    eng.sync()
    values_dev = []
    for s in range(W):
        plan.run_device_slot(slots[my_terms[s]])
        values_dev.append(plan.read_output()[0])
    eng.trace(True)                          # per-launch CUDA events inside the timed region (roofline of the dominant kernel)
    eng.read_trace()
    eng.reset_stats()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    eng.timer_start()
    for s in range(W, W + K):
        plan.run_device_slot(slots[my_terms[s]])
    ms_dev = eng.timer_stop()
    barrier()
    stats_dev = eng.stats()
    trace = eng.read_trace()
    eng.trace(False)
    values_dev.append(plan.read_output()[0])

    # ---------------- end-to-end path: files in, host mirror, scalar out ----------------------------------------
    values_e2e = []
    for s in range(W):
        v, fl, nodes, _ = host_api.contract_linegraph(QASM, meas_files[my_terms[s]], ORDERING, True)
        values_e2e.append(v)
    eng.reset_stats()
    barrier()
    # the K timed terms, one job in flight ahead of the one being read: files -> host mirror -> device -> scalar for every
    # term, with the host bookkeeping of term s+1 overlapping the device work of term s (host_api.LinegraphJob)
    eng.timer_start()
    job = host_api.LinegraphJob(QASM, meas_files[my_terms[W]], ORDERING, True)
    for s in range(W, W + K):
        nxt = host_api.LinegraphJob(QASM, meas_files[my_terms[s + 1]], ORDERING, True) if s + 1 < W + K else None
        v, fl, nodes = job.result()
        values_e2e.append(v)
        job = nxt
    ms_e2e = eng.timer_stop()
    barrier()
    sampler.stop()
    stats_e2e = eng.stats()

    # ---------------- second half of the metric: one amplitude index-sliced over the ranks --------------------------
    # The <Z27 Z29> term cut into 4^s slices (qtorch_b200/slicing.py: wires chosen greedily; two wires: peak rank 14 -> 12,
    # total units x1.005); slices dealt round-robin to ranks, one compiled plan, per-slice inputs staged in HBM, the
    # partial sums meet in one NCCL allreduce per amplitude.  Strong scaling of ONE expectation value.
    from qtorch_b200 import slicing
    from qtorch_b200.dispatch import Dispatcher
    golden_rec = json.load(open(NETS))["qaoa30_z27z29"]
    g_ranks, g_steps, g_inputs, _ = host_api.export_plan_linegraph(QASM, os.path.join(GOLDEN, golden_rec["measure"]), ORDERING, True)
    # as few slices as give every rank work: 4 slices (one wire, peak rank 13) up to 4 ranks, 16 slices (two wires, peak
    # rank 12) for 8 -- bigger slices keep the tile kernel's prologue and tail a smaller share of each step
    SLICE_WIRES = 1
    while 4 ** SLICE_WIRES < world:
        SLICE_WIRES += 1
    wires = slicing.choose_wires(g_ranks, g_steps, SLICE_WIRES)
    all_sl = slicing.all_slices(wires)
    plan_launches = plan.launches
    plan.destroy()                                  # give the unsliced plan's 13 GB back first
    # steps no cut wire reaches (270 of the 299 here) are hoisted into a prefix that runs once per amplitude
    splan, cuts, n_invariant = slicing.compile_sliced(eng, g_ranks, g_steps, wires)
    disp = Dispatcher(rank, world)
    owned = disp.owned(len(all_sl))
    for slot, u in enumerate(owned):
        splan.stage_inputs(slot, slicing.slice_inputs(g_inputs, g_ranks, cuts, wires, all_sl[u]))
    if dist is not None:
        uid = [eng.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(world, rank, uid[0])

    def one_amplitude():
        part = splan.run_slots(range(len(owned))) if owned else 0.0 + 0.0j      # prefix once, suffix per slice, one sync
        if dist is not None:
            part = complex(eng.allreduce_sum(np.array([part], dtype=np.complex128))[0])
        return part

    for _ in range(2):
        amp = one_amplitude()
    K2 = 5
    barrier()
    eng.timer_start()
    for _ in range(K2):
        amp = one_amplitude()
    ms_sliced = eng.timer_stop()
    barrier()
    sliced_ok = abs(amp - complex(*golden_rec["value"])) <= 1e-10
    sliced_units = splan.prefix_units + (splan.units - splan.prefix_units) * len(all_sl)
    splan.destroy()

    # ---------------- beyond the reference's plan: the same term on the in-process min-fill ordering -------------------
    # (LineGraph::runMinFill).  NOT the headline: the headline keeps the reference's own QuickBB plan bit
    # for bit; this shows what dropping the external quickbb_64 call buys (36 ms of ordering instead of 20 s, and a plan
    # of 2.2e10 instead of 6.9e10 units for this circuit).  Same value within 1e-10.
    m_ranks, m_steps, m_inputs, m_flops = host_api.export_plan_linegraph(QASM, os.path.join(GOLDEN, golden_rec["measure"]), "", True)
    mplan = eng.plan(m_ranks, m_steps)
    mplan.stage_inputs(0, m_inputs)
    for _ in range(3):
        mplan.run_device_slot(0)
    mval = complex(mplan.read_output()[0])
    barrier()
    eng.timer_start()
    for _ in range(10):
        mplan.run_device_slot(0)
    ms_minfill = eng.timer_stop() / 10
    minfill_ok = abs(mval - complex(*golden_rec["value"])) <= 1e-10
    mplan.destroy()

    # both paths must agree with each other (and with the golden term when it is among them)
    for s in range(W):
        assert abs(values_dev[s] - values_e2e[s]) <= 1e-10 * max(1.0, abs(values_e2e[s])), (s, values_dev[s], values_e2e[s])
    golden = json.load(open(NETS))["qaoa30_z27z29"]["value"]
    if (29, 27) in edges and my_terms[0] == edges.index((29, 27)):
        assert abs(values_e2e[0] - complex(*golden)) <= 1e-10

    # max over ranks of the timed regions; one scalar allreduce of the per-rank partial sums (the term dispatcher's reduction)
    f_p = sum(0.5 * (1.0 - v.real) for v in values_e2e[W:])
    if dist is not None:
        t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(t[0]), float(t[1])
        t2 = torch.tensor([ms_sliced], dtype=torch.float64, device="cuda")
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms_sliced = float(t2[0])
        f_p = eng.allreduce_sum(np.array([f_p], dtype=np.complex128))[0].real
    barrier()

    if rank == 0:
        terms = K * world
        value = terms / (ms_dev * 1e-3)
        e2e = terms / (ms_e2e * 1e-3)
        # roofline of the dominant kernel: the DMMA tile kernel on the rank-14 steps (trace code 2; the fourth rank-14 step
        # runs fused with the closing inner product, code 6, and is left out of this average)
        gett = [r for r in trace if r["kernel"] == 2 and max(r["rank_a"], r["rank_b"]) >= 9 and r["k"] == 3 and r["rank_a"] + r["rank_b"] - 6 == 14]
        total_ms = sum(r["ms"] for r in trace) or 1.0
        roof = None
        if gett:
            avg_ms = sum(r["ms"] for r in gett) / len(gett)
            flop = 8.0 * 4 ** 17                                   # 8 * 4^(rC+k), rC = 14, k = 3
            ach = flop / (avg_ms * 1e-3) / 1e12
            variant = os.environ.get("QTB_GETT_C1", "2")
            three_m = variant in ("2", "3")
            roof = {"bound": "tensor", "kernel": "k_gett (warp-specialised FP64 DMMA tiles, the four rank-14 steps), variant %s (%s complex product)" % (variant, "3M" if three_m else "4M"),
                    "achieved": ach, "peak": FP64_PEAK_TFLOPS,
                    "unit": "TFLOP/s", "frac": ach / FP64_PEAK_TFLOPS,
                    # achieved counts the ALGORITHMIC 8 flops per complex MAC; the 3M kernel issues 6 on the tensor pipe
                    "tensor_pipe_flops_per_unit": 6 if three_m else 8, "tensor_pipe_frac": ach * (0.75 if three_m else 1.0) / FP64_PEAK_TFLOPS,
                    # dram__bytes_read.sum + dram__bytes_write.sum of one (10,10,k=3 -> 14) launch, ncu --set full
                    # (profiles/r01_ncu_summary.txt: 0.259 GB read + 4.238 GB written); algorithmic 16 * (2 * 4^10 + 4^14) = 4.33e9
                    "traffic": 4.50e9, "launches_timed": len(gett), "avg_ms": avg_ms,
                    "share_of_step": sum(r["ms"] for r in gett) / total_ms,
                    # the fourth rank-14 step runs fused with the closing inner product (trace code 6, same tile kernel)
                    "fused_step_share_of_step": sum(r["ms"] for r in trace if r["kernel"] == 6) / total_ms,
                    "peak_source": "measured: tools/probe_fp64 DMMA m8n8k4 on this pool's B200 (profiles/r01_probe_fp64.jsonl); MEASURED_PEAKS.json has no FP64 entry"}
        by_kind = {}
        for r in trace:
            by_kind.setdefault(str(r["kernel"]), [0, 0.0])
            by_kind[str(r["kernel"])][0] += 1
            by_kind[str(r["kernel"])][1] += r["ms"]
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            c = cpu_reference_sample(threads)
            if c is not None:
                cpu = {"value": c["value"], "unit": "terms/s", "cores": threads, "kind": "reference",
                       "sample": "unmodified reference (-O2) replaying plan steps 0..%d of 299 (%.3g of %.3g units, %.1f s), extrapolated by units"
                                 % (c["steps"] - 1, c["units"], UNITS_PER_TERM, c["seconds"])}
        line = {
            "metric": METRIC, "value": value, "unit": "terms/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_dev / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "terms_per_step": world, "l2": "inputs larger than L2 (rank-14 tensors, 4.29 GB each)",
                       "plan_launches_per_term": plan_launches, "units_per_term": UNITS_PER_TERM},
            "e2e": {"value": e2e, "unit": "terms/s", "ms_per_step": ms_e2e / K,
                    "h2d_bytes_per_step": stats_e2e["bytes_h2d"] // K, "d2h_bytes_per_step": stats_e2e["bytes_d2h"] // K},
            "gpu_launches": int(stats_dev["launches"]),
            "roofline": roof, "cpu_baseline": cpu, "clocks": sampler.summary(),
            "kernel_time_ms_by_kind": {k: {"launches": v[0], "ms": v[1]} for k, v in by_kind.items()},
            "f_p_partial": f_p,
            "sliced": {"metric": "sliced_amplitudes_per_s", "value": K2 / (ms_sliced * 1e-3), "unit": "amplitudes/s", "ms_per_amplitude": ms_sliced / K2,
                       "workload": "cfg2 term <Z27 Z29> cut into 4^%d slices dealt round-robin over %d rank(s), one NCCL allreduce per amplitude" % (SLICE_WIRES, world),
                       "slices": len(all_sl), "invariant_steps_run_once": n_invariant, "peak_rank": slicing.plan_cost(g_ranks, g_steps, frozenset(wires))[1],
                       "units_vs_unsliced": sliced_units / UNITS_PER_TERM, "matches_reference_1e-10": bool(sliced_ok), "scaling": "strong"},
            "minfill_plan": {"note": "same term on the in-process min-fill ordering instead of the reference's QuickBB plan (not the headline)",
                             "terms_per_s_per_gpu": 1e3 / ms_minfill, "ms_per_term": ms_minfill, "units_per_term": m_flops,
                             "matches_reference_1e-10": bool(minfill_ok)},
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
