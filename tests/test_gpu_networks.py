"""GPU parity, network level: the C++ host mirror (Network / LineGraph / ContractionTools on the B200 engine) and
the compiled-plan path of the C ABI against the golden values produced by the unmodified reference
(tests/golden/networks.json; tolerance from BASELINE.json north_star: |d| <= 1e-10 * max(|ref|, 1))."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_paths, plan_file
import qtorch_b200 as qt
from oracle import oracle as O

pytestmark = pytest.mark.gpu
NETS = json.load(open(os.path.join(GOLDEN, "networks.json")))
SMALL = sorted(n for n in NETS if n != "qaoa30_z27z29")
TOL = 1e-10


def _close(val, ref):
    ref = complex(ref[0], ref[1])
    return abs(val - ref) <= TOL * max(abs(ref), 1.0)


def _run(name, tmp_path, extra_env=None, steps=False):
    rec = NETS[name]
    cwd, qasm, meas, ordering = golden_paths(rec)
    if rec["method"] == "lg":
        args = ["lg+steps" if steps else "lg", qasm, meas, ordering, rec["reduce"]]
    else:
        args = ["seq+steps" if steps else "seq", qasm, meas, plan_file(rec, tmp_path)]
    return rec, qt.run_harness(args, cwd=cwd, extra_env=extra_env, timeout=600)


@pytest.mark.parametrize("name", SMALL)
def test_network_value_matches_reference(built, name, tmp_path):
    rec, out = _run(name, tmp_path)
    # disconnected networks make the line-graph path fail in the reference too (ContractionFailure, partial value kept)
    assert out.get("exception") == rec.get("exception"), out.get("exception")
    val = complex(float(out["value"][0]), float(out["value"][1]))
    assert _close(val, rec["value"]), (val, rec["value"])
    assert out["plan"] == rec["plan"] and int(out["flops"][0]) == rec["flops"]


@pytest.mark.parametrize("name", ["qft8_X8", "qaoa20_node1_m125", "ghz64_zeros", "two_pairs_0011"])
def test_network_value_without_micro_grouping(built, name, tmp_path):
    """every step as its own launch (QTB_NO_MICRO) must give the same answer as the grouped executor"""
    rec, out = _run(name, tmp_path, extra_env={"QTB_NO_MICRO": "1"})
    val = complex(float(out["value"][0]), float(out["value"][1]))
    assert _close(val, rec["value"]), (val, rec["value"])


def test_plan_api_against_oracle(engine):
    """qtb_plan_*: a random tree of contractions (micro + big steps mixed), host inputs in, result out"""
    rng = np.random.default_rng(42)
    ranks = [3, 4, 2, 5, 6, 1, 7, 8]
    inputs = [rng.standard_normal(4 ** r) + 1j * rng.standard_normal(4 ** r) for r in ranks]
    # (a, b, posA, posB): ids >= len(ranks) are results
    steps = [
        (0, 1, [0, 2], [1, 3]),          # 3x4 k2 -> rank 3   id 8
        (2, 8, [1], [0]),                # 2x3 k1 -> rank 3   id 9
        (3, 4, [0, 4], [5, 2]),          # 5x6 k2 -> rank 7   id 10
        (5, 10, [0], [3]),               # 1x7 k1 -> rank 6   id 11
        (6, 7, [1, 3, 5], [0, 4, 7]),    # 7x8 k3 -> rank 9   id 12  (DMMA tiles)
        (11, 12, [0, 2, 4], [8, 1, 3]),  # 6x9 k3 -> rank 9   id 13
        (9, 13, [0, 1, 2], [2, 5, 7]),   # 3x9 k3 -> rank 6   id 14
    ]
    plan = engine.plan(ranks, steps)
    out = plan.run_host(inputs)
    tens = list(inputs)
    rk = list(ranks)
    O.lib().qto_set_threads(16)
    units = 0
    for a, b, pa, pb in steps:
        tens.append(O.contract(tens[a], rk[a], tens[b], rk[b], pa, pb))
        rk.append(rk[a] + rk[b] - 2 * len(pa))
        units += 4 ** (rk[-1] + len(pa))
    assert plan.output_rank == rk[-1] and plan.units == units
    assert np.abs(out - tens[-1]).max() <= 1e-11 * np.abs(tens[-1]).max()
    # device-resident replay gives the same bits every time (deterministic kernels)
    plan.upload_inputs(inputs)
    plan.run_device(); a1 = plan.read_output()
    plan.run_device(); a2 = plan.read_output()
    assert np.array_equal(a1.view(np.float64), a2.view(np.float64)) and np.array_equal(a1.view(np.float64), out.view(np.float64))
    plan.destroy()


def test_plan_rejects_reuse_of_contracted_tensor(engine):
    with pytest.raises(qt.EngineError):
        engine.plan([1, 1, 1], [(0, 1, [0], [0]), (0, 2, [0], [0])])


def test_config2_qaoa30_term_full_size(built, tmp_path):
    """BASELINE config 2 at full size: the <Z27 Z29> term of 4regRand30Node5-p1 on the frozen ordering --
    6.9e10 units, four rank-14 DMMA steps and a 268M-term inner product -- against the reference's value
    (-0.26965021147727669, 797 s on 8 CPU threads)."""
    rec, out = _run("qaoa30_z27z29", tmp_path)
    assert "exception" not in out, out.get("exception")
    val = complex(float(out["value"][0]), float(out["value"][1]))
    assert _close(val, rec["value"]), (val, rec["value"])
    assert out["plan"] == rec["plan"] and int(out["flops"][0]) == rec["flops"] == 69351174176


def test_config2_trace_is_one_full_size(built, tmp_path):
    """size-independent property on the same 30-qubit plan: measuring nothing (all qubits traced) must give
    tr(rho) = 1 -- exercises every big kernel with a different set of rank-1 caps"""
    rec = NETS["qaoa30_z27z29"]
    cwd, qasm, _, ordering = golden_paths(rec)
    meas = os.path.join(str(tmp_path), "allT.txt")
    open(meas, "w").write(" ".join(["T"] * 30) + "\n")
    out = qt.run_harness(["lg", qasm, meas, ordering, 1], cwd=cwd, timeout=600)
    val = complex(float(out["value"][0]), float(out["value"][1]))
    assert abs(val - 1.0) <= 1e-10, val
    assert out["plan"] == rec["plan"]


def test_reference_test_suite_drop_in(built, tmp_path):
    """the reference's own src/tests.cpp (13 known-answer tests: rotations, Bell/cat/teleportation/
    Toffoli, arbitrary 1- and 2-qubit gates vs its in-test state-vector simulator, large QAOA circuits, random circuits,
    disconnected networks, user-defined sequences, line graph + QuickBB), compiled UNMODIFIED against the B200 host
    mirror (oracle/_ref/tests_dropin, built where /root/reference exists), must pass 13/13 on the GPU."""
    import shutil
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "oracle", "_ref", "tests_dropin")
    qbb = os.path.join(ROOT, "oracle", "_ref", "quickbb_64")
    if not (os.path.exists(exe) and os.path.exists(qbb)):
        pytest.skip("oracle/_ref/tests_dropin not built (needs /root/reference at build time)")
    work = os.path.join(str(tmp_path), "run")
    os.makedirs(work)
    shutil.copytree(os.path.join(GOLDEN, "Samples"), os.path.join(work, "Samples"))
    os.makedirs(os.path.join(work, "output"))
    env = dict(os.environ)
    env["PATH"] = os.path.dirname(qbb) + ":" + env["PATH"]
    env["QTORCH_QUIET"] = "1"
    p = subprocess.run([exe], cwd=work, capture_output=True, text=True, timeout=1500, env=env)
    log = open(os.path.join(work, "output", "testingoutput.log")).read()
    lines = log.splitlines()
    assert "TOTAL TEST FAILURE COUNT: 0." in log, log[-3000:]
    assert sum(1 for l in lines if l.strip() == "Passed") == 13 and not any(l.strip() == "Failed" for l in lines), log[-3000:]
    assert p.returncode == 0, p.stdout[-2000:]


@pytest.mark.parametrize("name", ["qft8_X8", "testJW_YXXY", "qaoa20_node1_m125", "rand20_cn3_d12_zeros", "ghz64_zeros", "qaoa30_z27z29"])
def test_in_process_minfill_ordering_gives_the_reference_value(built, name, tmp_path):
    """LineGraph::runMinFill() replaces the external quickbb_64 call; a different (here cheaper) plan,
    the same value"""
    rec = NETS[name]
    cwd, qasm, meas, _ = golden_paths(rec)
    out = qt.run_harness(["minfill", qasm, meas, os.path.join(str(tmp_path), "mf.qbb.out"), rec["reduce"]], cwd=cwd, timeout=600)
    assert int(out["ok"][0]) == 1
    val = complex(float(out["value"][0]), float(out["value"][1]))
    assert _close(val, rec["value"]), (val, rec["value"])
    if name == "qaoa30_z27z29":
        assert int(out["flops"][0]) < rec["flops"]          # 2.2e10 vs QuickBB's 6.9e10 units


def test_linegraph_jobs_in_flight_match_the_blocking_call(built):
    """host_api.LinegraphJob (begin: parse + enqueue, no sync; result(): read back) with several networks in flight
    gives exactly what the blocking contract_linegraph gives, whatever order the results are collected in."""
    from qtorch_b200 import host_api
    names = ["qft8_X8", "qaoa20_node1_m125", "testJW_YXXY", "qft8_X8"]
    recs = [NETS[n] for n in names]
    paths = []
    for rec in recs:
        cwd, qasm, meas, ordering = golden_paths(rec)
        paths.append((os.path.join(cwd, qasm), meas if os.path.isabs(meas) else os.path.join(cwd, meas),
                      ordering if os.path.isabs(ordering) else os.path.join(cwd, ordering), bool(rec["reduce"])))
    blocking = [host_api.contract_linegraph(*p) for p in paths]
    jobs = [host_api.LinegraphJob(*p) for p in paths]
    for i in (2, 0, 3, 1):
        v, flops, nodes = jobs[i].result()
        assert abs(v - blocking[i][0]) <= 1e-14 and flops == blocking[i][1] == recs[i]["flops"] and nodes == blocking[i][2]
        assert _close(v, recs[i]["value"])
    with pytest.raises(RuntimeError):
        jobs[0].result()


COST = json.load(open(os.path.join(GOLDEN, "cost_plans.json")))


@pytest.mark.parametrize("name", sorted(COST))
def test_cost_based_planner_value(built, name):
    """the sampled greedy planner end to end on the device: the reference's plan (same seed) and the reference's value"""
    c = COST[name]
    cwd, qasm, meas, _ = golden_paths(NETS[c["network"]])
    out = qt.run_harness(["cost", qasm, meas, c["p"], c["seed"]], cwd=cwd, timeout=300)
    assert "exception" not in out, out.get("exception")
    val = complex(float(out["value"][0]), float(out["value"][1]))
    assert _close(val, c["value"]), (val, c["value"])
    assert out["plan"] == c["plan"] and int(out["flops"][0]) == c["flops"]


@pytest.mark.gpu
def test_plan_cache_replays_one_compiled_plan_for_many_measurements(built, networks):
    """host/PlanCache.h: the plan depends on (circuit topology, ordering), not on the measurement string -- the three
    test_JW observables share ONE compiled plan (first call compiles, the others are cache hits) and still match the
    reference's values, which it computed with a separate ordering per observable"""
    from qtorch_b200 import host_api
    first = networks["testJW_XXXX"]
    _, qasm, _, ordering = golden_paths(first)
    qasm = os.path.join(GOLDEN, first["qasm"])
    hits = []
    for name in ("testJW_XXXX", "testJW_YXXY", "testJW_Z0", "testJW_XXXX"):
        rec = networks[name]
        v, flops, nodes, hit = host_api.contract_cached(qasm, os.path.join(GOLDEN, rec["measure"]), ordering, True)
        ref = complex(*rec["value"])
        assert abs(v - ref) <= 1e-10 * max(1.0, abs(ref)), (name, v, ref)
        assert flops == first["flops"] and nodes == first["nodes"]
        hits.append(hit)
    assert hits == [False, True, True, True]
    # GHZ-1000: zeros and ones on one compiled plan
    z, o = networks["ghz1000_zeros"], networks["ghz1000_ones"]
    gq, go = os.path.join(GOLDEN, z["qasm"]), os.path.join(GOLDEN, z["ordering"])
    for rec in (z, o, z):
        v, flops, nodes, hit = host_api.contract_cached(gq, os.path.join(GOLDEN, rec["measure"]), go, True)
        assert abs(v - complex(*rec["value"])) <= 1e-10 and flops == rec["flops"]


@pytest.mark.parametrize("csize", [3, 8])
@pytest.mark.parametrize("name", ["qft8_X8", "qaoa20_node1_m125", "qaoa20_node5_m125", "ghz64_zeros", "two_pairs_0011", "rand20_cn3_d12_zeros"])
def test_grouped_micro_steps_on_thread_block_clusters(built, name, csize, tmp_path):
    """QTB_MICRO_CLUSTER=c forces EVERY grouped launch onto a thread-block cluster of c CTAs that share each level's items and meet
    in a cluster barrier between levels (kernels.cuh; by default only heavy groups take that path -- the micro-steps of a config-2
    term, p=2 QAOA terms).  Same value, same plan, same unit count as the reference."""
    rec, out = _run(name, tmp_path, extra_env={"QTB_MICRO_CLUSTER": str(csize)})
    val = complex(float(out["value"][0]), float(out["value"][1]))
    assert _close(val, rec["value"]), (val, rec["value"])
    assert out["plan"] == rec["plan"] and int(out["flops"][0]) == rec["flops"]
