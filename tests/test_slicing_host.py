"""The C++14 host's index-slicing planner (qtorch_b200/host/Slicing.h) against the Python planner the oracle tests use
(qtorch_b200/slicing.py): same wires, same sliced plan, same hoisting, same tensor slices.  GPU: the sliced-amplitude
executor (qtb_sliced_*: lanes, device-side accumulation, two amplitudes in flight) against the reference's values of the
UNSLICED networks."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_paths
from qtorch_b200 import host_api, slicing

NETS = json.load(open(os.path.join(GOLDEN, "networks.json")))


def _export(name):
    rec = NETS[name]
    cwd, qasm, meas, ordering = golden_paths(rec)
    return rec, host_api.export_plan_linegraph(os.path.join(cwd, qasm), meas, ordering, bool(rec["reduce"]))


@pytest.mark.parametrize("name,s", [("qft8_X8", 1), ("qft8_X8", 2), ("rand20_cn3_d12_zeros", 2), ("testJW_YXXY", 3), ("qaoa20_node5_m125", 2),
                                    ("qaoa30_z27z29", 2), ("ghz64_zeros", 1)])
def test_cpp_planner_equals_python_planner(built, name, s):
    rec, (ranks, steps, inputs, flops) = _export(name)
    wires = slicing.choose_wires(ranks, steps, s)
    ranks2, steps2, cuts = slicing.slice_plan(ranks, steps, wires)
    steps3, n_inv = slicing.hoist_invariant(len(ranks2), steps2, cuts.keys())
    got = host_api.slice_plan(ranks, steps, s)
    assert got["wires"] == [tuple(w) for w in wires]
    assert got["ranks"] == ranks2 and got["n_invariant"] == n_inv
    assert got["steps"] == [(a, b, list(pa), list(pb)) for a, b, pa, pb in steps3]
    assert {t: [(l, tuple(w)) for l, w in c] for t, c in cuts.items()} == got["cuts"]
    units1, peak1 = slicing.plan_cost(ranks, steps, frozenset(wires))
    assert got["units_per_slice"] == float(units1) and got["peak_rank"] == peak1


def test_cpp_tensor_slices_equal_numpy(built):
    rng = np.random.default_rng(3)
    for rank, legs, digits in [(3, [1], [2]), (4, [0, 3], [1, 3]), (2, [0, 1], [3, 0]), (5, [2], [0]), (1, [0], [2])]:
        x = rng.standard_normal(4 ** rank) + 1j * rng.standard_normal(4 ** rank)
        idx = [slice(None)] * rank
        for l, d in zip(legs, digits):
            idx[l] = d
        want = x.reshape((4,) * rank, order="F")[tuple(idx)].reshape(-1, order="F")
        assert np.array_equal(host_api.slice_tensor(x, rank, legs, digits), want)


def test_slices_are_dealt_round_robin():
    """the dispatcher's rule: slice u belongs to rank u % world; every slice has exactly one owner"""
    from qtorch_b200.dispatch import deal_round_robin
    for world in (1, 2, 4, 8):
        owned = [deal_round_robin(16, r, world) for r in range(world)]
        assert sorted(u for o in owned for u in o) == list(range(16))
        assert all(u % world == r for r, o in enumerate(owned) for u in o)


@pytest.mark.gpu
@pytest.mark.parametrize("name,s,lanes", [("qft8_X8", 2, 1), ("qft8_X8", 2, 3), ("rand20_cn3_d12_zeros", 2, 2), ("qaoa20_node1_m125", 1, 2),
                                          ("qaoa20_node1_m125", 2, 4), ("testJW_XXXX", 0, 2)])
def test_sliced_executor_matches_reference(engine, name, s, lanes):
    rec = NETS[name]
    cwd, qasm, meas, ordering = golden_paths(rec)
    net = host_api.SlicedNetwork(os.path.join(cwd, qasm), meas, ordering, bool(rec["reduce"]), slice_wires=s, lanes=lanes)
    ref = complex(*rec["value"])
    assert net.slices == 4 ** s and net.owned == net.slices and net.units_unsliced == rec["flops"]
    # one at a time, then two amplitudes in flight on alternating slot banks
    for _ in range(2):
        v = net.amplitude()
        assert abs(v - ref) <= 1e-10 * max(1.0, abs(ref)), (v, ref)
    net.stage(0); net.stage(1)
    tickets = [net.begin(0), net.begin(1)]
    for i in range(6):
        v = net.end(tickets[i % 2])
        assert abs(v - ref) <= 1e-10 * max(1.0, abs(ref)), (i, v, ref)
        tickets[i % 2] = net.begin(i % 2)
    for t in tickets:
        assert abs(net.end(t) - ref) <= 1e-10 * max(1.0, abs(ref))
    net.close()


@pytest.mark.gpu
def test_sliced_executor_is_deterministic_and_minfill(engine):
    """fixed accumulation order: repeated amplitudes are bit-identical; the in-process min-fill ordering gives the same value"""
    rec = NETS["qaoa20_node5_m125"]
    cwd, qasm, meas, ordering = golden_paths(rec)
    net = host_api.SlicedNetwork(os.path.join(cwd, qasm), meas, ordering, True, slice_wires=2, lanes=2)
    vals = [net.amplitude() for _ in range(4)]
    assert len({(v.real, v.imag) for v in vals}) == 1
    net.close()
    net2 = host_api.SlicedNetwork(os.path.join(cwd, qasm), meas, "", True, slice_wires=1, lanes=2)
    assert abs(net2.amplitude() - complex(*rec["value"])) <= 1e-10
    net2.close()
