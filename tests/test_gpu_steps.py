"""GPU parity, step level: every kernel family of libqtorch_b200 against the CPU oracle (the C restatement of
Network::ContractIndices, /root/reference/src/Network.h:876-971) on the same seeded inputs, called through the
C ABI.  FP64 complex; tolerance 1e-12 relative to max|C| (north_star allows 1e-10 on final values; a single step
differs from the reference only by FMA contraction and summation order)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
import qtorch_b200 as qt
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _rand(rank, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(4 ** rank) + 1j * rng.standard_normal(4 ** rank)


def _check(engine, rA, rB, pA, pB, seed=0):
    A, B = _rand(rA, seed), _rand(rB, seed + 1)
    ta, tb = engine.tensor(rA, A), engine.tensor(rB, B)
    tc = engine.contract(ta, tb, pA, pB)
    C = tc.download()
    O.lib().qto_set_threads(8)
    ref = O.contract(A, rA, B, rB, pA, pB)
    for t in (ta, tb, tc):
        t.free()
    err = np.abs(C - ref).max() / max(1.0, np.abs(ref).max())
    assert err <= TOL, (rA, rB, pA, pB, err)
    return err


def test_golden_steps_from_reference(engine):
    g = np.load(os.path.join(GOLDEN, "steps.npz"))
    n = len([k for k in g.files if k.startswith("spec")])
    for i in range(n):
        spec = g["spec%d" % i]
        rA, rB, k = (int(x) for x in spec[:3])
        pA, pB = [int(x) for x in spec[3:3 + k]], [int(x) for x in spec[3 + k:3 + 2 * k]]
        rng = np.random.default_rng(1000 + i)
        A = rng.standard_normal(4 ** rA) + 1j * rng.standard_normal(4 ** rA)
        B = rng.standard_normal(4 ** rB) + 1j * rng.standard_normal(4 ** rB)
        ta, tb = engine.tensor(rA, A), engine.tensor(rB, B)
        C = engine.contract(ta, tb, pA, pB).download()
        ref = g["C%d" % i]
        assert np.abs(C - ref).max() <= TOL * max(1.0, np.abs(ref).max()), i


# micro-step executor (U <= 4^8, ranks <= 7)
@pytest.mark.parametrize("rA,rB,pA,pB", [
    (1, 1, [0], [0]), (0, 0, [], []), (2, 1, [0], [0]), (4, 2, [3], [0]), (4, 4, [2, 3], [0, 1]), (3, 3, [0, 1, 2], [1, 2, 0]),
    (5, 3, [0, 4], [2, 1]), (4, 4, [1], [2]), (6, 2, [5], [1]), (2, 6, [0], [0]), (4, 4, [0, 1, 2, 3], [3, 1, 0, 2]), (7, 1, [3], [0]),
    (3, 2, [], []), (5, 5, [0, 1, 2, 3], [0, 1, 2, 3]),
])
def test_micro_steps(engine, rA, rB, pA, pB):
    _check(engine, rA, rB, pA, pB, seed=3)


def test_micro_dependency_chain(engine):
    """a chain of deferred micro-steps in ONE grouped launch: levels must order producers before consumers"""
    rng = np.random.default_rng(5)
    vecs = [rng.standard_normal(16) + 1j * rng.standard_normal(16) for _ in range(40)]     # rank-2 "gates"
    state = rng.standard_normal(4) + 1j * rng.standard_normal(4)
    engine.sync()
    before = engine.stats()
    t = engine.tensor(1, state)
    ref = state.copy()
    for g in vecs:
        tg = engine.tensor(2, g)
        t2 = engine.contract(t, tg, [0], [0])
        t.free(); tg.free()
        t = t2
        ref = O.contract(ref, 1, g, 2, [0], [0])
    out = t.download()
    after = engine.stats()
    assert np.abs(out - ref).max() <= 1e-10 * np.abs(ref).max()
    assert after["steps"] - before["steps"] == 40
    assert after["launches"] - before["launches"] <= 2, "40 tiny steps must be grouped, not launched one by one"


# generic thread / warp kernels
@pytest.mark.parametrize("rA,rB,pA,pB", [
    (6, 6, [0, 5], [3, 1]),            # rC 8, thread kernel
    (5, 5, [4], [4]),                  # rC 8, K 4
    (7, 6, [0, 6], [0, 5]),            # rC 9
    (8, 8, [0, 1, 2, 3, 4, 5], [5, 4, 3, 2, 1, 0]),   # rC 4, K 4096: warp kernel
    (8, 7, [1, 2, 3, 4, 5], [0, 2, 4, 5, 6]),         # rC 5, K 1024
    (9, 9, [0, 1, 2, 3, 4, 5, 6, 7, 8], [8, 7, 6, 5, 4, 3, 2, 1, 0]),   # rank 0, K 4^9: split-K reduce
    (9, 8, [1, 2, 3, 4, 5, 6, 7, 8], [7, 0, 6, 1, 5, 2, 4, 3]),         # rC 1, split-K with 4 outputs
    (8, 8, [0, 1, 2, 3, 4, 5, 6], [6, 5, 4, 3, 2, 1, 0]),               # rC 2, split-K with 16 outputs
])
def test_generic_and_reduce_steps(engine, rA, rB, pA, pB):
    _check(engine, rA, rB, pA, pB, seed=7)


# tiled DMMA kernel: every tile configuration, both operand roles, scattered leg positions
@pytest.mark.parametrize("rA,rB,pA,pB", [
    (7, 7, [0, 2, 5], [1, 6, 3]),       # C1/TK16: M=N=256, K=64
    (7, 7, [1, 4], [0, 3]),             # C1/TK16: K=16, M=N=1024
    (8, 4, [7], [0]),                   # C1/TK4: M=4^7, N=64, K=4
    (4, 8, [0], [7]),                   # swapped roles
    (8, 3, [2], [1]),                   # C2/TK4: N=16
    (8, 4, [0, 5], [3, 1]),             # C2/TK16: N=16, K=16
    (8, 2, [3], [0]),                   # C3/TK4: N=4 padded to 8
    (8, 1, [6], [0]),                   # C3/TK4: N=1
    (9, 3, [0, 4, 8], [2, 0, 1]),       # C3/TK16: N=1, K=64
    (9, 4, [1, 3, 7], [0, 3, 2]),       # C3/TK16: N=4, K=64
    (3, 9, [0, 1, 2], [8, 4, 0]),       # swapped, N=1
    (6, 9, [0, 2, 3], [2, 7, 8]),       # the config-2 "gate apply" shape scaled down: M=64, K=64, N=4^6
    (9, 6, [1, 6, 8], [0, 1, 4]),       # same with roles exchanged
    (8, 8, [0, 2, 5, 6], [1, 7, 6, 0]), # K=256: four k-chunks
    (10, 2, [9], [1]),                  # streaming: rank-10 x gate
])
def test_gett_steps(engine, rA, rB, pA, pB):
    _check(engine, rA, rB, pA, pB, seed=11)


def test_gett_matches_generic_kernels_exactly_enough(engine):
    """same step through the DMMA tiles and through the plain FMA kernel (QTB_FORCE_GENERIC is process-wide, so
    compare against the oracle with a tighter bound instead): relative 1e-13 on O(1) data"""
    err = _check(engine, 7, 7, [0, 3, 6], [6, 3, 0], seed=13)
    assert err < 1e-13


def test_linearity_property_large(engine):
    """size-independent property at a size the oracle cannot reach in seconds (rank-11 result):
    contract(A, B1 + 2*B2) == contract(A, B1) + 2*contract(A, B2)"""
    rA, rB, pA, pB = 9, 8, [0, 5, 7], [3, 2, 4]
    A, B1, B2 = _rand(rA, 1), _rand(rB, 2), _rand(rB, 3)
    ta = engine.tensor(rA, A)
    outs = []
    for B in (B1, B2, B1 + 2 * B2):
        tb = engine.tensor(rB, B)
        tc = engine.contract(ta, tb, pA, pB)
        outs.append(tc.download())
        tb.free(); tc.free()
    # contract() consumed nothing: A is still valid
    lhs, rhs = outs[2], outs[0] + 2 * outs[1]
    assert np.abs(lhs - rhs).max() <= 1e-11 * np.abs(rhs).max()
    O.lib().qto_set_threads(16)
    ref = O.contract(A, rA, B1, rB, pA, pB)
    assert np.abs(outs[0] - ref).max() <= TOL * np.abs(ref).max()
    ta.free()


def test_error_statuses(engine):
    ta, tb = engine.tensor(2, _rand(2, 1)), engine.tensor(2, _rand(2, 2))
    with pytest.raises(qt.EngineError) as e:
        engine.contract(ta, tb, [1, 0], [0, 1])          # pos_a not increasing
    assert e.value.status == 2
    with pytest.raises(qt.EngineError):
        engine.contract(ta, tb, [0], [5])
    empty = engine.tensor(2)                               # never uploaded: the reference throws InvalidFunctionInput
    with pytest.raises(qt.EngineError) as e:
        engine.contract(ta, empty, [0], [0])
    assert e.value.status == 3
    wrong = engine.tensor(3)
    with pytest.raises(qt.EngineError):
        engine.contract(ta, tb, [0], [0], out=wrong)      # rank(C) must be rA + rB - 2k
    # the engine is still usable afterwards
    c = engine.contract(ta, tb, [0], [0])
    ref = O.contract(ta.download(), 2, tb.download(), 2, [0], [0])
    assert np.abs(c.download() - ref).max() < 1e-13


def test_upload_download_roundtrip_all_small_ranks(engine):
    for r in range(0, 9):
        x = _rand(r, 100 + r)
        t = engine.tensor(r, x)
        assert np.array_equal(t.download(), x)
        if r == 0:
            assert t.scalar() == x[0]
        t.free()
