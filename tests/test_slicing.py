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
    assert info["slices"] == 4 ** s
    assert abs(total - ref) <= 1e-10 * max(1.0, abs(ref)), (total, ref)
