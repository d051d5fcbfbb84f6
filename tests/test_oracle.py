"""The oracle itself: C restatement of Network::ContractIndices (/root/reference/src/Network.h:876-971)
pinned against (a) golden single-step answers produced by the unmodified reference, (b) an independent numpy
restatement, (c) the live reference binary when oracle/_ref is present."""
import os
import tempfile

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import oracle as O


def _rand(rank, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(4 ** rank) + 1j * rng.standard_normal(4 ** rank)


def test_oracle_matches_reference_golden_steps_bit_exact(built):
    g = np.load(os.path.join(GOLDEN, "steps.npz"))
    n = len([k for k in g.files if k.startswith("spec")])
    assert n >= 20
    for i in range(n):
        spec = g["spec%d" % i]
        rA, rB, k = (int(x) for x in spec[:3])
        pA, pB = [int(x) for x in spec[3:3 + k]], [int(x) for x in spec[3 + k:3 + 2 * k]]
        rng = np.random.default_rng(1000 + i)
        A = rng.standard_normal(4 ** rA) + 1j * rng.standard_normal(4 ** rA)
        B = rng.standard_normal(4 ** rB) + 1j * rng.standard_normal(4 ** rB)
        C = O.contract(A, rA, B, rB, pA, pB)
        assert np.array_equal(C.view(np.float64), g["C%d" % i].view(np.float64)), "step %d not bit-exact" % i
        assert O.step_units(rA, rB, k) == int(g["flops%d" % i][0])


@pytest.mark.parametrize("rA,rB,pA,pB", [(3, 4, [0, 2], [3, 1]), (6, 3, [1, 5], [2, 0]), (5, 5, [0, 1, 2, 3, 4], [4, 3, 2, 1, 0]),
                                          (0, 0, [], []), (7, 2, [3], [1]), (2, 7, [0, 1], [6, 2]), (4, 4, [], [])])
def test_oracle_vs_numpy(built, rA, rB, pA, pB):
    A, B = _rand(rA, 1), _rand(rB, 2)
    C = O.contract(A, rA, B, rB, pA, pB)
    Cn = O.contract_numpy(A, rA, B, rB, pA, pB)
    assert C.shape == Cn.shape
    assert np.abs(C - Cn).max() <= 1e-12 * max(1.0, np.abs(Cn).max())


def test_oracle_thread_count_does_not_change_bits(built):
    A, B = _rand(7, 3), _rand(6, 4)
    O.lib().qto_set_threads(1)
    C1 = O.contract(A, 7, B, 6, [0, 3, 6], [5, 1, 2])
    O.lib().qto_set_threads(5)
    C5 = O.contract(A, 7, B, 6, [0, 3, 6], [5, 1, 2])
    O.lib().qto_set_threads(1)
    assert np.array_equal(C1.view(np.float64), C5.view(np.float64))


def test_oracle_rejects_bad_leg_maps(built):
    A, B = _rand(2, 1), _rand(2, 2)
    with pytest.raises(ValueError):
        O.contract(A, 2, B, 2, [1, 0], [0, 1])       # pos_a must be increasing (A-wire order, Network.h:739-749)
    with pytest.raises(ValueError):
        O.contract(A, 2, B, 2, [0], [2])


def test_final_value_rule(built):
    L = O.lib()
    assert L.qto_final_value_rule(0.0, 0.0, 3, 3) == 1           # first scalar result always taken
    assert L.qto_final_value_rule(0.5, 0.0, 3, 3) == 0           # later component results do not overwrite ...
    assert L.qto_final_value_rule(0.5, 0.0, 0, 0) == 1           # ... but a product of two scalars does (Network.h:963)
    assert L.qto_final_value_rule(1e-31, -1e-31, 2, 1) == 1


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_vs_live_reference(built):
    for seed, (rA, rB, pA, pB) in enumerate([(4, 3, [1, 2], [0, 2]), (2, 5, [0], [3]), (6, 6, [0, 2, 4], [5, 3, 1])]):
        A, B = _rand(rA, 10 + seed), _rand(rB, 20 + seed)
        with tempfile.TemporaryDirectory() as d:
            Cr, out = O.ref_step(A, rA, B, rB, pA, pB, d)
        C = O.contract(A, rA, B, rB, pA, pB)
        assert np.array_equal(C.view(np.float64), Cr.view(np.float64))


def test_shim_patcher_replaces_exactly_one_function_body():
    """oracle/make_shim.py (the minimal shim of INTEGRATION.md) cuts the body of Network::ContractIndices out of the reference header
    by brace matching: nested blocks inside the body go, the declaration in the class and every other function stay."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_shim", os.path.join(os.path.dirname(O.__file__), "make_shim.py"))
    ms = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ms)
    header = (
        "#pragma once\n#include <vector>\nclass Network {\n    inline void ContractIndices(int a, std::shared_ptr<Node> nodeC);\n};\n"
        "inline void Network::Other() { if (x) { y(); } }\n"
        "    inline void Network::ContractIndices(int a,\n            std::shared_ptr<Node> nodeC) {\n"
        "        auto f1 = [&]() { for (;;) { if (a) { break; } } };\n        if (a) { f1(); } else { g(); }\n    }\n"
        "inline void Network::After() { z(); }\n")
    out = ms.patched_network_h(header, "        BODY();")
    assert out.count("BODY();") == 1 and "f1" not in out and "g();" not in out
    assert "inline void Network::Other() { if (x) { y(); } }" in out and "inline void Network::After() { z(); }" in out
    assert "inline void ContractIndices(int a, std::shared_ptr<Node> nodeC);" in out            # the in-class declaration is untouched
    assert out.index('#include "qtorch_b200.h"') < out.index("#include <vector>")
    assert out.count("{") == out.count("}")
