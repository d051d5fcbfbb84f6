#!/usr/bin/env python
"""tools/bench_configs.py -- the BASELINE configs that are parity cases rather than the bench line (1, 3, 4-literal, 5),
timed end to end through the host mirror on one B200.  One JSON object per line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["QTORCH_QUIET"] = "1"
from qtorch_b200 import host_api  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
NETS = json.load(open(os.path.join(G, "networks.json")))
MC = json.load(open(os.path.join(G, "maxcut.json")))
eng = host_api.engine()


def time_network(name, reps=5):
    rec = NETS[name]
    qasm, meas, ordering = (os.path.join(G, rec[k]) for k in ("qasm", "measure", "ordering"))
    best, val = 1e9, None
    for _ in range(reps):
        eng.reset_stats()
        t0 = time.perf_counter()
        val, flops, nodes, secs = host_api.contract_linegraph(qasm, meas, ordering, bool(rec["reduce"]))
        dt = time.perf_counter() - t0
        best = min(best, secs)
    st = eng.stats()
    ok = abs(val - complex(*rec["value"])) <= 1e-10 * max(1.0, abs(complex(*rec["value"])))
    print(json.dumps({"config": name, "value": [val.real, val.imag], "matches_reference": bool(ok), "steps": st["steps"], "launches": st["launches"],
                      "contraction_ms_after_parse": best * 1e3, "steps_per_s": st["steps"] / best, "units": flops, "wall_ms_incl_parse": dt * 1e3}))


def time_maxcut(name, evals=200):
    rec = MC[name]
    t0 = time.perf_counter()
    q = host_api.QaoaObjective(os.path.join(G, rec["graph"]), rec["p"])
    plan_s = time.perf_counter() - t0
    vals, fp = q.evaluate(rec["betas_gammas"])
    ok = abs(fp - rec["fp"]) <= 1e-10 * rec["fp"]
    eng.reset_stats()
    t0 = time.perf_counter()
    bg = list(rec["betas_gammas"])
    for i in range(evals):
        bg[0] = rec["betas_gammas"][0] + 1e-3 * (i % 7)
        q.evaluate(bg)
    dt = time.perf_counter() - t0
    st = eng.stats()
    print(json.dumps({"config": "maxcutQAOA " + name, "fp": fp, "matches_reference": bool(ok), "edges": len(vals), "evaluations": evals,
                      "ms_per_evaluation": dt / evals * 1e3, "terms_per_s": evals * len(vals) / dt, "launches_per_evaluation": st["launches"] / evals,
                      "units_per_evaluation": q.units, "planning_s": plan_s}))
    q.close()


def time_plan_replay(name, reps=50):
    """the same network as a compiled plan replayed on resident inputs (one CUDA graph / grouped launches): device time only"""
    rec = NETS[name]
    qasm, meas, ordering = (os.path.join(G, rec[k]) for k in ("qasm", "measure", "ordering"))
    ranks, steps, inputs, flops = host_api.export_plan_linegraph(qasm, meas, ordering, bool(rec["reduce"]))
    plan = eng.plan(ranks, steps)
    plan.upload_inputs(inputs)
    for _ in range(3):
        plan.run_device()
    val = complex(plan.read_output()[0])
    eng.timer_start()
    for _ in range(reps):
        plan.run_device()
    ms = eng.timer_stop() / reps
    ok = abs(val - complex(*rec["value"])) <= 1e-10 * max(1.0, abs(complex(*rec["value"])))
    print(json.dumps({"config": name + " (compiled plan replay)", "matches_reference": bool(ok), "steps": len(steps), "launches": plan.launches,
                      "ms_per_replay": ms, "steps_per_s": len(steps) / (ms * 1e-3), "units": flops}))
    plan.destroy()


def time_cached(name, reps=20):
    """the same call through the plan cache (host/PlanCache.h): parse + reduce + ordering walk + plan compilation once, then one
    small H2D (the caps), one graph launch and one 16-byte D2H per evaluation"""
    rec = NETS[name]
    qasm, meas, ordering = (os.path.join(G, rec[k]) for k in ("qasm", "measure", "ordering"))
    t0 = time.perf_counter()
    val, flops, nodes, hit = host_api.contract_cached(qasm, meas, ordering, bool(rec["reduce"]))
    first = time.perf_counter() - t0
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        val, flops, nodes, hit = host_api.contract_cached(qasm, meas, ordering, bool(rec["reduce"]))
        best = min(best, time.perf_counter() - t0)
    ok = abs(val - complex(*rec["value"])) <= 1e-10 * max(1.0, abs(complex(*rec["value"])))
    print(json.dumps({"config": name + " (plan cache)", "matches_reference": bool(ok), "first_call_ms": first * 1e3, "cached_call_ms": best * 1e3,
                      "cache_hit": bool(hit), "units": flops}))


time_network("qft8_X8")
time_plan_replay("qft8_X8")
time_plan_replay("ghz1000_zeros")
time_network("ghz1000_zeros")
time_network("ghz1000_ones")
time_cached("ghz1000_zeros")
time_cached("ghz1000_ones")
time_cached("qft8_X8")
time_network("qaoa20_node5_m125")
time_network("rand20_cn3_d12_zeros")
time_maxcut("3reg30_p1_default", 300)
time_maxcut("4reg30_p1_default", 300)
time_maxcut("3reg30_p2_default", 30)
