"""Host logic (no GPU): the C++ host mirror must emit EXACTLY the reference's contraction plan -- node ids,
step order, mCreatedFrom pairs, getNumFloatOps -- for every golden case (SURVEY.md 8a plan-identity contract),
the same lg.cnf text, and the same gate tensors.  Runs the mirror in plan-only mode (no arithmetic)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_paths, plan_file
import qtorch_b200 as qt
from oracle import oracle as O

NETS = json.load(open(os.path.join(GOLDEN, "networks.json")))
LG_CASES = sorted(n for n, r in NETS.items() if r["method"] == "lg")
STOCH_CASES = sorted(n for n, r in NETS.items() if r["method"] == "stoch")


@pytest.mark.parametrize("name", LG_CASES)
def test_linegraph_plan_identical_to_reference(built, name):
    rec = NETS[name]
    cwd, qasm, meas, ordering = golden_paths(rec)
    out = qt.run_harness(["lg", qasm, meas, ordering, rec["reduce"]], plan_only=True, cwd=cwd)
    assert int(out["ok"][0]) == rec["ok"]
    assert out.get("exception") == rec.get("exception")
    assert out["plan"] == rec["plan"]
    assert int(out["flops"][0]) == rec["flops"]
    assert int(out["nodes"][0]) == rec["nodes"]


@pytest.mark.parametrize("name", STOCH_CASES)
def test_recorded_stochastic_plan_replays(built, name, tmp_path):
    """a plan found by the reference's stochastic search replays step for step (ContractGivenSequence)"""
    rec = NETS[name]
    cwd, qasm, meas, _ = golden_paths(rec)
    out = qt.run_harness(["seq", qasm, meas, plan_file(rec, tmp_path)], plan_only=True, cwd=cwd)
    assert int(out["done"][0]) == 1
    assert out["plan"] == rec["plan"]
    assert int(out["flops"][0]) == rec["flops"]
    assert int(out["nodes"][0]) == rec["nodes"]


def test_stochastic_search_completes_and_is_seedable(built):
    rec = NETS["rand6_rxyz_stoch"]
    cwd, qasm, meas, _ = golden_paths(rec)
    a = qt.run_harness(["stoch", qasm, meas, 7], plan_only=True, cwd=cwd)
    b = qt.run_harness(["stoch", qasm, meas, 7], plan_only=True, cwd=cwd)
    c = qt.run_harness(["stoch", qasm, meas, 8], plan_only=True, cwd=cwd)
    assert a["plan"] == b["plan"] and int(a["nodes"][0]) == rec["nodes"]
    assert a["plan"] != c["plan"]


COST = json.load(open(os.path.join(GOLDEN, "cost_plans.json")))


@pytest.mark.parametrize("name", sorted(COST))
def test_cost_based_planner_draws_the_reference_plan(built, name):
    """ContractionTools::Contract(CostContractSimple, p) is a randomised search; with the generator seeded like the
    reference run that made the golden file it must take the same decisions draw by draw: identical plan and float ops"""
    c = COST[name]
    cwd, qasm, meas, _ = golden_paths(NETS[c["network"]])
    out = qt.run_harness(["cost", qasm, meas, c["p"], c["seed"]], plan_only=True, cwd=cwd, timeout=120)
    assert "exception" not in out, out.get("exception")
    assert out["plan"] == c["plan"]
    assert int(out["flops"][0]) == c["flops"] and int(out["nodes"][0]) == c["nodes"]


def test_cost_based_planner_lookahead_two_terminates(built):
    """pValue = 2: the reference redraws forever once a super-node has absorbed its whole neighbourhood
    (ContractionTools.h:1002-1004); the mirror stops looking ahead there and still finishes with a complete plan"""
    rec = NETS["qft8_X8"]
    cwd, qasm, meas, _ = golden_paths(rec)
    out = qt.run_harness(["cost", qasm, meas, 2, 1], plan_only=True, cwd=cwd, timeout=120)
    assert "exception" not in out and int(out["nodes"][0]) == rec["nodes"] and len(out["plan"]) == len(rec["plan"])


def test_bell_pair_worked_example(built):
    """the smallest plan KAT -- (0,2)->6, (3,5)->7, (7,4)->8, (6,8)->9, (1,9)->done; 356 units; 11 nodes"""
    rec = NETS["bell_00"]
    assert rec["plan"] == ["0,2", "3,5", "7,4", "6,8", "1,9"] and rec["flops"] == 356 and rec["nodes"] == 11


@pytest.mark.skipif(not O.ref_available() or not os.path.exists("/root/reference/bin/quickbb_64"), reason="needs the reference tree")
@pytest.mark.parametrize("name", ["qft8_X8", "testJW_XXXX", "rand6_rxyz_YI"])
def test_cnf_text_identical_to_reference(built, name, tmp_path):
    rec = NETS[name]
    cwd, qasm, meas, _ = golden_paths(rec)
    mine = os.path.join(str(tmp_path), "mine.cnf")
    qt.run_harness(["cnf", qasm, meas, mine, rec["reduce"]], plan_only=True, cwd=cwd)
    env = dict(os.environ)
    env["PATH"] = "/root/reference/bin:" + env["PATH"]
    ref = os.path.join(str(tmp_path), "ref.cnf")
    O.ref_harness(["qbb", qasm, meas, 1, ref, os.path.join(str(tmp_path), "o"), os.path.join(str(tmp_path), "s"), rec["reduce"]], cwd=cwd, env=env)
    assert open(mine).read() == open(ref).read()


def test_gate_tensors_match_reference_tables(built):
    gates = json.load(open(os.path.join(GOLDEN, "gates.json")))
    assert len(gates) >= 40
    for key, rec in gates.items():
        parts = key.split()
        args = [os.path.join(GOLDEN, parts[1])] if parts[0].startswith("def") else parts[1:]
        out = qt.run_harness(["gate", parts[0]] + args, plan_only=True)
        assert int(out["gate"][1]) == rec["rank"], key
        mine = np.array([float(x) for x in out["gate"][2:]])
        ref = np.array(rec["vals"])
        assert mine.shape == ref.shape, key
        assert np.abs(mine - ref).max() <= 4e-16, (key, np.abs(mine - ref).max())


def test_parser_error_behaviour(built, tmp_path):
    bad = os.path.join(str(tmp_path), "bad.qasm")
    meas = os.path.join(str(tmp_path), "m.txt")
    open(meas, "w").write("0 0\n")
    open(bad, "w").write("2\nH 0\nFOO 1\n")
    out = qt.run_harness(["stoch", bad, meas], plan_only=True)
    assert out["exception"] == "Invalid File Format."
    open(bad, "w").write("2\nCNOT 0 2\n")
    assert qt.run_harness(["stoch", bad, meas], plan_only=True)["exception"] == "Invalid File Format."
    out = qt.run_harness(["stoch", os.path.join(str(tmp_path), "missing.qasm"), meas], plan_only=True)
    assert out["exception"] == "Invalid Input or Output File Path"


@pytest.mark.parametrize("name", ["qft8_X8", "testJW_YXXY", "rand6_rxz_ZI", "cat8_ones", "qaoa20_node5_m125", "ghz64_zeros"])
def test_exported_plan_evaluated_by_oracle_matches_reference_value(built, name):
    """host mirror (parser, gate tensors, ReduceCircuit, LGContract bookkeeping) + CPU oracle arithmetic reproduce the
    reference's value: pins everything except the CUDA kernels without a GPU"""
    from qtorch_b200 import host_api
    rec = NETS[name]
    cwd, qasm, meas, ordering = golden_paths(rec)
    ranks, steps, inputs, flops = host_api.export_plan_linegraph(os.path.join(cwd, qasm), meas, ordering, bool(rec["reduce"]))
    assert flops == rec["flops"] and len(steps) == len(rec["plan"])
    O.lib().qto_set_threads(8)
    t, rk = list(inputs), list(ranks)
    for a, b, pa, pb in steps:
        t.append(O.contract(t[a], rk[a], t[b], rk[b], pa, pb))
        rk.append(rk[a] + rk[b] - 2 * len(pa))
    ref = complex(*rec["value"])
    assert abs(complex(t[-1][0]) - ref) <= 1e-12 * max(1.0, abs(ref))


def test_inp_script_reader_matches_reference(built):
    """the ".inp" reader (host/leviParser.hpp): the four maps it fills from a script with comments, every accepted bool
    spelling class, a bad bool, an unknown type and a directive after free text -- against the dump the reference's
    leviParser produced for the same file (tests/golden/scripts/reader_cases.expected), and live when it is built"""
    script = os.path.join(GOLDEN, "scripts", "reader_cases.inp")
    want = open(os.path.join(GOLDEN, "scripts", "reader_cases.expected")).read().split("\n")
    want = [l for l in want if l]
    import subprocess
    got = subprocess.run([qt.HARNESS_PATH, "inp", script], capture_output=True, text=True, timeout=60,
                         env=dict(os.environ, QTORCH_PLAN_ONLY="1", QTORCH_QUIET="1")).stdout.splitlines()
    assert [l for l in got if l.startswith("@@")] == want
    if O.ref_available():
        ref = subprocess.run([O.ref_harness_path(), "inp", script], capture_output=True, text=True, timeout=60).stdout.splitlines()
        assert [l for l in ref if l.startswith("@@")] == want
    missing = subprocess.run([qt.HARNESS_PATH, "inp", os.path.join(GOLDEN, "scripts", "no_such_file.inp")], capture_output=True, text=True,
                             timeout=60, env=dict(os.environ, QTORCH_PLAN_ONLY="1", QTORCH_QUIET="1")).stdout
    assert "@@opened 0" in missing
