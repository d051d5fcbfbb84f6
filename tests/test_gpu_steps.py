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


def _random_step(rng, max_units_log4):
    while True:
        rA, rB = int(rng.integers(0, 11)), int(rng.integers(0, 11))
        k = int(rng.integers(0, min(rA, rB) + 1))
        rC = rA + rB - 2 * k
        if rC > 11 or rC + k > max_units_log4 or (k == 0 and rC > 6):
            continue
        pA = sorted(rng.choice(rA, size=k, replace=False).tolist()) if k else []
        pB = rng.permutation(rB)[:k].tolist() if k else []
        return rA, rB, pA, pB


def test_random_leg_maps_all_kernel_families(engine):
    """200 random (rank, shared-leg) configurations up to 4^11 units: whatever kernel family the dispatcher picks
    (grouped micro-steps, thread/warp kernels, every DMMA tile shape with either operand role, split-K) must agree
    with the oracle; deferred micro-steps are flushed in batches, so grouping across unrelated steps is covered too"""
    rng = np.random.default_rng(2024)
    O.lib().qto_set_threads(16)
    pending = []
    for trial in range(200):
        rA, rB, pA, pB = _random_step(rng, 11)
        A, B = _rand(rA, 5000 + trial), _rand(rB, 7000 + trial)
        ta, tb = engine.tensor(rA, A), engine.tensor(rB, B)
        tc = engine.contract(ta, tb, pA, pB)
        pending.append((rA, rB, pA, pB, A, B, ta, tb, tc))
        if len(pending) == 8 or trial == 199:
            for (rA_, rB_, pA_, pB_, A_, B_, ta_, tb_, tc_) in pending:
                C = tc_.download()
                ref = O.contract(A_, rA_, B_, rB_, pA_, pB_)
                err = np.abs(C - ref).max() / max(1.0, np.abs(ref).max())
                assert err <= TOL, (rA_, rB_, pA_, pB_, err)
                for t in (ta_, tb_, tc_):
                    t.free()
            pending = []


def test_rank15_result_property(engine):
    """a 17 GB rank-15 result (beyond what the CPU oracle can check in seconds): gate application on a rank-15 tensor,
    checked through a size-independent property -- applying a unitary superoperator and then its inverse restores
    the tensor -- plus spot values against the definition"""
    rng = np.random.default_rng(7)
    r = 15
    A = None
    try:
        x = rng.standard_normal(4 ** 7) + 1j * rng.standard_normal(4 ** 7)          # rank-7 seed, expanded by outer products
        t7a, t7b = engine.tensor(7, x), engine.tensor(7, x[::-1].copy())
        t14 = engine.contract(t7a, t7b, [], [])                                       # rank 14 outer product (k = 0)
        v = rng.standard_normal(4) + 1j * rng.standard_normal(4)
        t1 = engine.tensor(1, v)
        A = engine.contract(t14, t1, [], [])                                          # rank 15: 17.2 GB
        for t in (t7a, t7b, t14, t1):
            t.free()
        # a one-qubit rotation superoperator S (rank 2) and its inverse on leg 9
        th = 0.7
        c, s = np.cos(th / 2), np.sin(th / 2)
        U = np.array([[c, -1j * s], [-1j * s, c]])
        def superop(U):
            S = np.zeros((4, 4), dtype=complex)
            for i in range(4):
                for o in range(4):
                    S[i, o] = U[o >> 1, i >> 1] * np.conj(U[o & 1, i & 1])
            return S.reshape(-1, order="F")
        g, ginv = engine.tensor(2, superop(U)), engine.tensor(2, superop(U.conj().T))
        B = engine.contract(A, g, [9], [0])            # legs: A's free legs (0..8, 10..14) then the gate's output leg
        C = engine.contract(B, ginv, [14], [0])        # undo it: back to the original values, same leg order as B
        probe = engine.tensor(1, np.array([1, 0, 0, 0], dtype=complex))
        # compare through two cheap reductions instead of downloading 17 GB: contract leg 14 of C and of the permuted A
        ca = engine.contract(C, probe, [14], [0])      # rank 14: C[..., 0]
        a9 = engine.contract(A, probe, [9], [0])       # rank 14: A[leg9 = 0]  (same remaining leg order)
        da, db = ca.download(), a9.download()
        assert np.abs(da - db).max() <= 1e-12 * np.abs(db).max()
        for t in (B, C, g, ginv, probe, ca, a9):
            t.free()
    finally:
        if A is not None:
            A.free()


@pytest.mark.parametrize("t_is_a", [True, False])
@pytest.mark.parametrize("shape", [(7, 7, [0, 3], [5, 1]), (6, 8, [1, 2], [7, 0]), (8, 6, [0, 4], [2, 5])])
def test_fused_dmma_step_plus_inner_product(engine, t_is_a, shape):
    """a big DMMA step whose rank-10 result is immediately contracted with another rank-10 tensor over all legs runs as
    ONE fused kernel (the intermediate never reaches HBM): eager path (held-back step) and compiled plan, both operand
    orders, against the oracle's two separate steps"""
    rA, rB, pA, pB = shape
    A, B = _rand(rA, 31), _rand(rB, 32)
    rT = rA + rB - 2 * len(pA)
    assert rT == 10
    D = _rand(rT, 33)
    perm = np.random.default_rng(5).permutation(rT).tolist()           # leg i of T pairs with leg perm[i] of D
    O.lib().qto_set_threads(16)
    T = O.contract(A, rA, B, rB, pA, pB)
    if t_is_a:
        ref = O.contract(T, rT, D, rT, list(range(rT)), perm)
        posA2, posB2 = list(range(rT)), perm
    else:
        order = np.argsort(perm).tolist()                               # D is operand A: its legs in increasing order
        posA2, posB2 = list(range(rT)), order                           # leg j of D pairs with leg order[j] of T
        ref = O.contract(D, rT, T, rT, posA2, posB2)
    # eager path
    before = engine.stats()
    ta, tb, td = engine.tensor(rA, A), engine.tensor(rB, B), engine.tensor(rT, D)
    tt = engine.contract(ta, tb, pA, pB)
    out = engine.contract(tt, td, posA2, posB2) if t_is_a else engine.contract(td, tt, posA2, posB2)
    val = out.scalar()
    after = engine.stats()
    assert abs(val - ref[0]) <= 1e-11 * max(1.0, abs(ref[0])), (val, ref[0])
    assert after["launches"] - before["launches"] <= 3                  # uploads + fused kernel + partial reduction
    for t in (ta, tb, td, tt, out):
        t.free()
    # compiled plan: inputs 0,1,2 = A, B, D; step 0 -> id 3 (T); step 1 -> id 4
    steps = [(0, 1, pA, pB), ((3, 2, posA2, posB2) if t_is_a else (2, 3, posA2, posB2))]
    plan = engine.plan([rA, rB, rT], steps)
    got = plan.run_host([A, B, D])[0]
    assert plan.launches == 2 and abs(got - ref[0]) <= 1e-11 * max(1.0, abs(ref[0]))
    plan.destroy()


@pytest.mark.parametrize("rA,rB,pA,pB", [
    (9, 1, [0], [0]),                   # measurement cap on leg 0: K=4, N=1, hole at the bottom
    (9, 1, [8], [0]),                   # ... on the top leg
    (9, 1, [4], [0]),                   # ... in the middle
    (1, 9, [0], [5]),                   # roles exchanged (the cap is the reference's node A)
    (9, 2, [3], [1]),                   # 1-qubit gate: K=4, N=4
    (2, 9, [0], [7]),                   # same, gate first: C = y + 4 x
    (9, 2, [0, 1], [1, 0]),             # K=16, N=1, both holes at the bottom (four lanes per output), pairs crossed
    (9, 3, [0, 1], [0, 2]),             # ... with N=4: stays with the tile kernel
    (3, 9, [1, 2], [1, 0]),             # ... N=4, small operand first: tile kernel
    (9, 2, [0, 5], [1, 0]),             # K=16, N=1: pairs crossed
    (9, 2, [2, 7], [0, 1]),             # K=16, N=1
    (9, 3, [0, 8], [2, 0]),             # K=16, N=4
    (3, 9, [0, 2], [8, 3]),             # K=16, N=4, small operand first, its shared legs out of order in the big one
    (9, 3, [4], [1]),                   # rank-3 tensor sharing one leg: K=4, N=16, big operand first
    (9, 3, [0], [2]),                   # ... shared leg = the big operand's leg 0 (one 64-byte run per index: 256-bit loads)
    (3, 9, [1], [6]),                   # K=4, N=16, small operand first: tile kernel (transposed 256-bit stores)
    (9, 2, [0], [0]),                   # K=4, N=4 with the shared leg at the bottom: 256-bit loads
    (2, 9, [1], [0]),                   # ... small operand first: 256-bit loads and 256-bit stores
    (8, 1, [], []),                     # outer product with a rank-1 tensor: K=1, N=4
    (1, 8, [], []),                     # outer product, small first: C = y + 4 x
    (10, 0, [], []),                    # scaling by a scalar tensor: K=1, N=1
])
def test_streaming_apply_steps(engine, rA, rB, pA, pB):
    """big tensor x tiny tensor (<= 64 elements): the one-thread-per-free-index streaming kernel (apply.cuh)"""
    engine.trace(True)
    engine.read_trace()
    _check(engine, rA, rB, pA, pB, seed=61)
    kinds = [t["kernel"] for t in engine.read_trace() if t["kernel"] != 0]       # (uploads ride in a grouped launch: code 0)
    engine.trace(False)
    if os.environ.get("QTB_NO_APPLY") != "1":
        tile = (len(pA) == 2 and min(rA, rB) == 3 and sorted(pA if rA > rB else pB) == [0, 1]) or (len(pA) == 1 and rA == 3 and rB > 3)
        assert kinds == [2 if tile else 7], kinds


def test_tile_kernels_without_the_streaming_class():
    """QTB_NO_APPLY=1 sends the gate-application shapes back through the DMMA tile kernel (k_gett 256x16 / 256x8 tiles,
    TK = 4): those instantiations must stay correct.  The switch is read once per process, hence the child."""
    import subprocess
    import sys
    env = dict(os.environ, QTB_NO_APPLY="1", QTORCH_QUIET="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k",
                        "gett_steps or random_leg_maps or streaming_apply_steps"],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:]


def test_deferred_steps_respect_reuse_of_buffers(engine):
    """Deferred micro-steps are levelled by dependency inside ONE grouped launch.  Rewriting a buffer that pending steps
    still read or write -- re-uploading a gate's angles into a live handle, re-using out= -- must be ordered after them
    (write-after-read / write-after-write), not scheduled side by side at level 0."""
    rng = np.random.default_rng(77)
    s0 = rng.standard_normal(4) + 1j * rng.standard_normal(4)
    g1, g2 = (rng.standard_normal(16) + 1j * rng.standard_normal(16) for _ in range(2))
    ts, tg = engine.tensor(1, s0), engine.tensor(2, g1)
    c1 = engine.contract(ts, tg, [0], [0])                 # pending: reads tg (= g1)
    tg.upload(g2)                                          # same handle, new contents, while c1 is still deferred
    c2 = engine.contract(ts, tg, [0], [0])
    r1, r2 = c1.download(), c2.download()
    assert np.abs(r1 - O.contract(s0, 1, g1, 2, [0], [0])).max() < 1e-13
    assert np.abs(r2 - O.contract(s0, 1, g2, 2, [0], [0])).max() < 1e-13
    # out= re-used: the second write lands after the step that reads the first result
    out = engine.tensor(1)
    engine.contract(ts, tg, [0], [0], out=out)             # out = s0 . g2
    d1 = engine.contract(out, tg, [0], [0])                # reads out
    tg2 = engine.tensor(2, g1)
    engine.contract(ts, tg2, [0], [0], out=out)            # rewrites out while d1 is pending
    want_d1 = O.contract(O.contract(s0, 1, g2, 2, [0], [0]), 1, g2, 2, [0], [0])
    assert np.abs(d1.download() - want_d1).max() < 1e-13
    assert np.abs(out.download() - O.contract(s0, 1, g1, 2, [0], [0])).max() < 1e-13
    for t in (ts, tg, tg2, c1, c2, out, d1):
        t.free()


def test_fused_intermediate_is_single_use(engine):
    """the eager fused path never writes the big intermediate: its handle must say so instead of serving pool garbage;
    self-contraction is rejected like in plans"""
    A, B, D = _rand(7, 41), _rand(7, 42), _rand(10, 43)
    ta, tb, td = engine.tensor(7, A), engine.tensor(7, B), engine.tensor(10, D)
    tt = engine.contract(ta, tb, [0, 3], [5, 1])
    out = engine.contract(tt, td, list(range(10)), list(range(10)))
    O.lib().qto_set_threads(8)
    ref = O.contract(O.contract(A, 7, B, 7, [0, 3], [5, 1]), 10, D, 10, list(range(10)), list(range(10)))[0]
    assert abs(out.scalar() - ref) <= 1e-11 * max(1.0, abs(ref))
    with pytest.raises(qt.EngineError) as e:
        tt.download()
    assert e.value.status == 3
    with pytest.raises(qt.EngineError) as e:
        engine.contract(tt, td, list(range(10)), list(range(10)))
    assert e.value.status == 3
    with pytest.raises(qt.EngineError) as e:
        engine.contract(ta, ta, [0], [1])
    assert e.value.status == 2
    for t in (ta, tb, td, tt, out):
        t.free()


def test_tma_fed_tile_kernel_matches_the_oracle():
    """QTB_TMA=1: compute-bound steps whose operand tiles qualify run on k_gett_tma (gett_tma.cuh: one cp.async.bulk.tensor
    box per operand and ring stage, 64-byte swizzle, host-searched fragment roles).  Opt-in because it measured slower than
    the cp.async gather; it must still be right, and it must actually be taken (tma_launches).  The switch is read once per
    process, hence the child."""
    import subprocess
    import sys
    code = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np
import qtorch_b200 as qt
from oracle import oracle as O
eng = qt.Engine(0)
rng = np.random.default_rng(11)
O.lib().qto_set_threads(8)
taken = 0
for rA, rB, pA, pB in [(8, 8, [0, 2, 5], [1, 7, 6]), (6, 10, [0, 2, 3], [2, 7, 9]), (8, 6, [0, 4], [2, 5]), (7, 9, [0, 4, 6], [6, 8, 5]),
                       (10, 6, [2, 5, 9], [0, 2, 3]), (7, 8, [0, 1], [0, 3]), (9, 7, [1, 4, 5], [5, 2, 3])]:
    A = rng.standard_normal(4 ** rA) + 1j * rng.standard_normal(4 ** rA)
    B = rng.standard_normal(4 ** rB) + 1j * rng.standard_normal(4 ** rB)
    eng.reset_stats()
    C = eng.contract(eng.tensor(rA, A), eng.tensor(rB, B), pA, pB).download()
    taken += eng.stats()["tma_launches"]
    ref = O.contract(A, rA, B, rB, pA, pB)
    assert np.abs(C - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), (rA, rB, pA, pB)
assert taken >= 3, taken
print("tma steps", taken)
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=dict(os.environ, QTB_TMA="1", QTORCH_QUIET="1"))
    assert r.returncode == 0 and "tma steps" in r.stdout, (r.stdout[-500:], r.stderr[-1500:])


@pytest.mark.gpu
def test_shared_sum_tile_kernel_matches_the_oracle():
    """QTB_GETT_C1=5: the compute-bound class on k_gett3s (gett3m.cuh: XOR-swizzled unpadded operand tiles, the 3M operand sums
    formed once per tile by the math warps in turn and shared through `summed` mbarriers), plain and fused with the inner
    product that follows, both ring depths.  Opt-in because it measured slower than k_gett (profiles/r02_g3_shared_sums.txt);
    it must still be right.  The switch is read once per process, hence the child."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for g3 in ("0", "1"):
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "check_g3.py"), "--parity-only"], capture_output=True, text=True, timeout=600,
                           env=dict(os.environ, QTB_GETT_C1="5", QTB_G3=g3, QTORCH_QUIET="1"))
        assert r.returncode == 0 and "parity failures: 0" in r.stdout and "WRONG" not in r.stdout, (r.stdout[-800:], r.stderr[-1500:])


@pytest.mark.gpu
def test_programmatic_dependent_launch_of_the_tile_kernel():
    """QTB_PDL=1: k_gett launches carry cudaLaunchAttributeProgrammaticStreamSerialization and the kernel's griddepcontrol.wait holds
    its first global access back until the previous grid has completed (gett.cuh).  Opt-in (measured neutral); back-to-back dependent
    tile-kernel steps -- plain and fused -- must still match the oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "check_g3.py"), "--parity-only"], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, QTB_PDL="1", QTORCH_QUIET="1"))
    assert r.returncode == 0 and "parity failures: 0" in r.stdout and "WRONG" not in r.stdout, (r.stdout[-800:], r.stderr[-1500:])
