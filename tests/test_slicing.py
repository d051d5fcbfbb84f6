"""Index slicing (BASELINE config 4): fixing s wires gives 4^s sub-plans with identical step shapes whose values sum to
the unsliced network value.  CPU: the sliced plans are evaluated with the oracle and compared with the reference's
value of the UNSLICED network; GPU: the same through one compiled device plan with staged input slots."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_paths
from qtorch_b200 import host_api, slicing
from oracle import oracle as O

NETS = json.load(open(os.path.join(GOLDEN, "networks.json")))


def _export(name):
    rec = NETS[name]
    cwd, qasm, meas, ordering = golden_paths(rec)
    return rec, host_api.export_plan_linegraph(os.path.join(cwd, qasm), meas, ordering, bool(rec["reduce"]))


def _oracle_run(ranks, steps, inputs):
    t, rk = list(inputs), list(ranks)
    for a, b, pa, pb in steps:
        t.append(O.contract(t[a], rk[a], t[b], rk[b], pa, pb))        # k = 0 (outer product) is legal for the oracle and the engine
        rk.append(rk[a] + rk[b] - 2 * len(pa))
    return complex(np.asarray(t[-1]).reshape(-1)[0])


@pytest.mark.parametrize("name,s", [("qft8_X8", 1), ("qft8_X8", 2), ("rand20_cn3_d12_zeros", 2), ("testJW_YXXY", 3), ("qaoa20_node5_m125", 1)])
def test_sliced_sum_equals_reference_value(built, name, s):
    rec, (ranks, steps, inputs, flops) = _export(name)
    wires = slicing.choose_wires(ranks, steps, s)
    assert len(wires) == s
    ranks2, steps2, cuts = slicing.slice_plan(ranks, steps, wires)
    units0, peak0 = slicing.plan_cost(ranks, steps)
    units1, peak1 = slicing.plan_cost(ranks, steps, frozenset(wires))
    assert units0 == flops and peak1 <= peak0 and units1 < units0
    O.lib().qto_set_threads(8)
    total = 0
    for digits in slicing.all_slices(wires):
        total += _oracle_run(ranks2, steps2, slicing.slice_inputs(inputs, ranks, cuts, wires, digits))
    ref = complex(*rec["value"])
    assert abs(total - ref) <= 1e-11 * max(1.0, abs(ref))


@pytest.mark.parametrize("name,s", [("qft8_X8", 2), ("rand20_cn3_d12_zeros", 2), ("testJW_YXXY", 1)])
def test_hoisting_keeps_every_slice_bit_identical(built, name, s):
    """hoist_invariant only swaps independent steps: same pairwise contractions, so each slice value is bit-identical,
    and no step of the invariant prefix may (transitively) read a cut input."""
    rec, (ranks, steps, inputs, flops) = _export(name)
    wires = slicing.choose_wires(ranks, steps, s)
    ranks2, steps2, cuts = slicing.slice_plan(ranks, steps, wires)
    steps3, n_inv = slicing.hoist_invariant(len(ranks2), steps2, cuts.keys())
    assert len(steps3) == len(steps2) and 0 < n_inv < len(steps2)
    tainted = set(cuts.keys())
    for i, (a, b, pa, pb) in enumerate(steps3):
        assert a < len(ranks2) + i and b < len(ranks2) + i
        if a in tainted or b in tainted:
            assert i >= n_inv
            tainted.add(len(ranks2) + i)
    assert len(ranks2) + len(steps3) - 1 in tainted          # the root depends on every cut
    O.lib().qto_set_threads(8)
    for digits in slicing.all_slices(wires)[:3]:
        ins = slicing.slice_inputs(inputs, ranks, cuts, wires, digits)
        assert _oracle_run(ranks2, steps2, ins) == _oracle_run(ranks2, steps3, ins)
    # nothing varies -> everything is invariant
    assert slicing.hoist_invariant(len(ranks), steps, [])[1] == len(steps)


def test_slicing_reduces_peak_rank(built):
    rec, (ranks, steps, inputs, flops) = _export("qaoa20_node1_m125")
    _, peak0 = slicing.plan_cost(ranks, steps)
    wires = slicing.choose_wires(ranks, steps, 3)
    _, peak3 = slicing.plan_cost(ranks, steps, frozenset(wires))
    assert peak3 < peak0


@pytest.mark.gpu
@pytest.mark.parametrize("name,s", [("qft8_X8", 2), ("rand20_cn3_d12_zeros", 2), ("qaoa20_node1_m125", 2), ("rand42_cn2_d20_zeros", 1)])
def test_sliced_contraction_on_device(engine, name, s):
    rec = NETS[name]
    if rec["method"] != "lg":
        pytest.skip("needs a line-graph plan export")
    rec, (ranks, steps, inputs, flops) = _export(name)
    wires = slicing.choose_wires(ranks, steps, s)
    total, info = slicing.contract_sliced(engine, ranks, steps, inputs, wires)
    ref = complex(*rec["value"])
    assert info["slices"] == 4 ** s and info["invariant_steps"] > 0
    assert abs(total - ref) <= 1e-10 * max(1.0, abs(ref)), (total, ref)


@pytest.mark.gpu
def test_run_slots_matches_slot_by_slot(engine):
    """qtb_plan_run_slots (invariant prefix once + suffix per slot, one sync) against the plain per-slot runs of the
    un-hoisted plan, slice by slice, twice in a row (the kept-alive prefix results must survive a full pass)."""
    rec, (ranks, steps, inputs, flops) = _export("qaoa20_node1_m125")
    wires = slicing.choose_wires(ranks, steps, 2)
    ranks2, steps2, cuts = slicing.slice_plan(ranks, steps, wires)
    plain = engine.plan(ranks2, steps2)
    hoisted, cuts_h, n_inv = slicing.compile_sliced(engine, ranks, steps, wires)
    assert cuts_h == cuts and 0 < n_inv < len(steps2) and 0 < hoisted.prefix_units < hoisted.units == plain.units
    sl = slicing.all_slices(wires)
    for slot, digits in enumerate(sl):
        ins = slicing.slice_inputs(inputs, ranks, cuts, wires, digits)
        plain.stage_inputs(slot, ins)
        hoisted.stage_inputs(slot, ins)
    want = []
    for slot in range(len(sl)):
        plain.run_device_slot(slot)
        want.append(complex(plain.read_output()[0]))
    for _ in range(2):
        total, each = hoisted.run_slots(range(len(sl)), each=True)
        assert np.allclose(each, np.array(want), rtol=1e-13, atol=1e-16)
        assert abs(total - complex(*rec["value"])) <= 1e-10
    # a subset in another order, and on the un-hoisted plan (no prefix: whole plan per slot)
    total, each = hoisted.run_slots([5, 2], each=True)
    assert np.allclose(each, [want[5], want[2]], rtol=1e-13, atol=1e-16)
    total, each = plain.run_slots([3], each=True)
    assert np.allclose(each[0], want[3], rtol=1e-13, atol=1e-16)
    plain.destroy()
    hoisted.destroy()
