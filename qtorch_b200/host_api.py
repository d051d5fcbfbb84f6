"""ctypes face of libqtorch_host.so: the C++ host mirror (Network / LineGraph / ContractionTools) called in-process.

``contract_linegraph`` is "the call a user makes" -- the flow of qtorch's main.cpp (/root/reference/src/main.cpp:74-198):
Network(qasm, measure) -> ReduceCircuit -> LineGraph(net).LGContract() on a frozen QuickBB ordering -> GetFinalValue().
``export_plan_linegraph`` runs the same host bookkeeping without the device and returns the plan + input tensors in the
layout qtb_plan_create expects.
"""
import ctypes
import os

import numpy as np

from . import PlanStep, load_library, DeviceUnavailable, Engine, QTB_MAX_RANK

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "libqtorch_host.so")
_hlib = None


class _QthPlan(ctypes.Structure):
    _fields_ = [("nInputs", ctypes.c_int), ("nSteps", ctypes.c_int), ("inputRanks", ctypes.POINTER(ctypes.c_int)),
                ("inputData", ctypes.POINTER(ctypes.c_double)), ("inputOffsets", ctypes.POINTER(ctypes.c_longlong)),
                ("steps", ctypes.POINTER(PlanStep)), ("flops", ctypes.c_longlong)]


class _QthSlicedPlan(ctypes.Structure):
    _fields_ = [("nInputs", ctypes.c_int), ("nSteps", ctypes.c_int), ("nInvariant", ctypes.c_int), ("nWires", ctypes.c_int),
                ("peakRank", ctypes.c_int), ("nCuts", ctypes.c_int), ("inputRanks", ctypes.POINTER(ctypes.c_int)),
                ("steps", ctypes.POINTER(PlanStep)), ("wires", ctypes.POINTER(ctypes.c_int)), ("cuts", ctypes.POINTER(ctypes.c_int)),
                ("unitsPerSlice", ctypes.c_double), ("unitsInvariant", ctypes.c_double)]


def host_library():
    global _hlib
    if _hlib is None:
        load_library()                                     # libqtorch_b200.so first (rpath covers it too)
        if not os.path.exists(HOST_LIB_PATH):
            raise DeviceUnavailable("libqtorch_host.so is not built (run `python -m qtorch_b200.build`)")
        os.environ.setdefault("QTORCH_QUIET", "1")
        H = ctypes.CDLL(HOST_LIB_PATH)
        H.qth_last_error.restype = ctypes.c_char_p
        H.qth_engine_ctx.restype = ctypes.c_void_p
        cd, cll, ci = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_int)
        H.qth_contract_linegraph.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, cd, cll, ci, cd]
        H.qth_contract_cached.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, cd, cll, ci, ci]
        H.qth_linegraph_begin.restype = ctypes.c_void_p
        H.qth_linegraph_begin.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
        H.qth_linegraph_end.argtypes = [ctypes.c_void_p, cd, cll, ci]
        H.qth_contract_sequence.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ci, ctypes.c_int, cd, cll, ci, cd]
        H.qth_export_plan_linegraph.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(_QthPlan)]
        H.qth_slice_plan.argtypes = [ctypes.c_int, ci, ctypes.c_int, ctypes.POINTER(PlanStep), ctypes.c_int, ctypes.POINTER(_QthSlicedPlan)]
        H.qth_slice_tensor.argtypes = [ctypes.c_void_p, ctypes.c_int, ci, ci, ctypes.c_int, ctypes.c_void_p]
        H.qth_sliced_create.restype = ctypes.c_void_p
        H.qth_sliced_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p] + [ctypes.c_int] * 5
        H.qth_sliced_destroy.argtypes = [ctypes.c_void_p]
        H.qth_sliced_info.argtypes = [ctypes.c_void_p, cll, cd, cd, cll]
        H.qth_sliced_stage.argtypes = [ctypes.c_void_p, ctypes.c_int]
        H.qth_sliced_begin.argtypes = [ctypes.c_void_p, ctypes.c_int]
        H.qth_sliced_end.argtypes = [ctypes.c_void_p, ctypes.c_int, cd]
        H.qth_maxcut_circuit_text.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, cd, ctypes.c_char_p, ctypes.c_int, ci, ci]
        H.qth_maxcut_final_string.argtypes = [ctypes.c_char_p, ctypes.c_int, cd, ctypes.c_char_p, ci, ctypes.c_int, ctypes.c_uint, cd]
        H.qth_qaoa_create.restype = ctypes.c_void_p
        H.qth_qaoa_create.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        H.qth_qaoa_destroy.argtypes = [ctypes.c_void_p]
        H.qth_qaoa_num_owned.argtypes = [ctypes.c_void_p]
        H.qth_qaoa_owned_edges.argtypes = [ctypes.c_void_p, ci]
        H.qth_qaoa_units.restype = ctypes.c_longlong
        H.qth_qaoa_units.argtypes = [ctypes.c_void_p]
        H.qth_qaoa_launches.argtypes = [ctypes.c_void_p]
        H.qth_qaoa_evaluate.argtypes = [ctypes.c_void_p, cd, ctypes.c_int, cd, cd]
        H.qth_qaoa_objective.argtypes = [ctypes.c_void_p, cd, ctypes.c_int, ctypes.c_int, ctypes.c_int, cd]
        H.qth_qaoa_circuit_text.argtypes = [ctypes.c_void_p, ctypes.c_int, cd, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
        _hlib = H
    return _hlib


def maxcut_circuit_text(graph_file, p, edge, betas_gammas):
    """host-only: (circuit text, number of edges, qubits of this term) exactly as the reference's F_p writes it"""
    H = host_library()
    bg = (ctypes.c_double * len(betas_gammas))(*betas_gammas)
    buf = ctypes.create_string_buffer(1 << 16)
    ne, nq = ctypes.c_int(), ctypes.c_int()
    n = H.qth_maxcut_circuit_text(graph_file.encode(), p, edge, bg, buf, len(buf), ctypes.byref(ne), ctypes.byref(nq))
    if n < 0:
        _raise(H, n)
    return buf.value.decode(), ne.value, nq.value


def maxcut_final_string(graph_file, p, betas_gammas, out_file, seed=1):
    """final cut string for the given angles (counterpart of the reference's maxcutGetFinalString)"""
    H = host_library()
    bg = (ctypes.c_double * len(betas_gammas))(*betas_gammas)
    bits = (ctypes.c_int * 4096)()
    prob = ctypes.c_double()
    n = H.qth_maxcut_final_string(graph_file.encode(), p, bg, out_file.encode(), bits, 4096, seed, ctypes.byref(prob))
    if n < 0:
        _raise(H, n)
    return [bits[i] for i in range(n)], prob.value


class QaoaObjective:
    """The B200 term dispatcher of host/maxcut.h: per-edge light-cone networks of a MaxCut QAOA instance, planned once,
    evaluated in one grouped launch per call.  rank/world select this process's share of the edges (edge % world)."""

    def __init__(self, graph_file, p=1, rank=0, world=1, plan_tries=8):
        self.H = host_library()
        self.p = p
        self.h = self.H.qth_qaoa_create(graph_file.encode(), p, rank, world, plan_tries)
        if not self.h:
            _raise(self.H, 1)
        n = self.H.qth_qaoa_num_owned(self.h)
        buf = (ctypes.c_int * max(n, 1))()
        self.H.qth_qaoa_owned_edges(self.h, buf)
        self.owned = [buf[i] for i in range(n)]
        self.units = self.H.qth_qaoa_units(self.h)
        self.launches = self.H.qth_qaoa_launches(self.h)

    def evaluate(self, betas_gammas):
        """-> (array of <ZZ> for the owned edges, this rank's partial F_p)"""
        bg = (ctypes.c_double * len(betas_gammas))(*betas_gammas)
        out = (ctypes.c_double * max(2 * len(self.owned), 2))()
        fp = ctypes.c_double()
        rc = self.H.qth_qaoa_evaluate(self.h, bg, len(betas_gammas), out, ctypes.byref(fp))
        if rc != 0:
            _raise(self.H, rc)
        vals = np.array([complex(out[2 * i], out[2 * i + 1]) for i in range(len(self.owned))])
        return vals, fp.value

    def objective(self, betas_gammas, reduce=False, n_edges_total=0):
        """F_p for the angles: one graph launch; reduce=True adds ONE in-stream NCCL allreduce over the job's ranks"""
        bg = (ctypes.c_double * len(betas_gammas))(*betas_gammas)
        fp = ctypes.c_double()
        rc = self.H.qth_qaoa_objective(self.h, bg, len(betas_gammas), 1 if reduce else 0, n_edges_total, ctypes.byref(fp))
        if rc != 0:
            _raise(self.H, rc)
        return fp.value

    def circuit_text(self, edge, betas_gammas):
        bg = (ctypes.c_double * len(betas_gammas))(*betas_gammas)
        buf = ctypes.create_string_buffer(1 << 16)
        self.H.qth_qaoa_circuit_text(self.h, edge, bg, len(betas_gammas), buf, len(buf))
        return buf.value.decode()

    def close(self):
        if self.h:
            self.H.qth_qaoa_destroy(self.h)
            self.h = None


def _raise(H, rc):
    msg = H.qth_last_error().decode()
    if "device engine unavailable" in msg:
        raise DeviceUnavailable(msg)
    raise RuntimeError("host mirror failed (rc=%d): %s" % (rc, msg))


def engine():
    """The host mirror's process-wide engine context, wrapped for stats / timers / trace."""
    H = host_library()
    ctx = H.qth_engine_ctx()
    if not ctx:
        _raise(H, 1)
    return Engine(ctx=ctx)


def contract_linegraph(qasm, measure, ordering, reduce=True):
    H = host_library()
    v = (ctypes.c_double * 2)()
    flops, nodes, secs = ctypes.c_longlong(), ctypes.c_int(), ctypes.c_double()
    rc = H.qth_contract_linegraph(qasm.encode(), measure.encode(), ordering.encode(), 1 if reduce else 0, v,
                                  ctypes.byref(flops), ctypes.byref(nodes), ctypes.byref(secs))
    if rc != 0:
        _raise(H, rc)
    return complex(v[0], v[1]), flops.value, nodes.value, secs.value


def contract_cached(qasm, measure, ordering, reduce=True):
    """contract_linegraph through the plan cache (host/PlanCache.h): host bookkeeping and plan compilation happen once per
    (qasm, ordering, reduce); later calls swap the measurement caps and replay the graph.  -> (value, flops, nodes, hit)"""
    H = host_library()
    v = (ctypes.c_double * 2)()
    flops, nodes, hit = ctypes.c_longlong(), ctypes.c_int(), ctypes.c_int()
    rc = H.qth_contract_cached(qasm.encode(), measure.encode(), (ordering or "").encode(), 1 if reduce else 0, v,
                               ctypes.byref(flops), ctypes.byref(nodes), ctypes.byref(hit))
    if rc != 0:
        _raise(H, rc)
    return complex(v[0], v[1]), flops.value, nodes.value, bool(hit.value)


class LinegraphJob:
    """contract_linegraph split in two: the constructor parses, reduces, walks the ordering and enqueues every step on
    the device without synchronising; ``result()`` reads the scalar back.  Starting job i+1 before asking for the
    result of job i hides the host bookkeeping of one network behind the device work of the previous one."""

    def __init__(self, qasm, measure, ordering, reduce=True):
        self.H = host_library()
        self.h = self.H.qth_linegraph_begin(qasm.encode(), measure.encode(), ordering.encode(), 1 if reduce else 0)
        if not self.h:
            _raise(self.H, 1)

    def result(self):
        """-> (value, float ops, nodes); the job is consumed"""
        if not self.h:
            raise RuntimeError("job already consumed")
        v = (ctypes.c_double * 2)()
        flops, nodes = ctypes.c_longlong(), ctypes.c_int()
        h, self.h = self.h, None
        rc = self.H.qth_linegraph_end(h, v, ctypes.byref(flops), ctypes.byref(nodes))
        if rc != 0:
            _raise(self.H, rc)
        return complex(v[0], v[1]), flops.value, nodes.value

    def __del__(self):
        if getattr(self, "h", None):
            v = (ctypes.c_double * 2)()
            self.H.qth_linegraph_end(self.h, v, None, None)
            self.h = None


def contract_sequence(qasm, measure, pairs):
    H = host_library()
    flat = (ctypes.c_int * (2 * len(pairs)))(*[x for p in pairs for x in p])
    v = (ctypes.c_double * 2)()
    flops, nodes, secs = ctypes.c_longlong(), ctypes.c_int(), ctypes.c_double()
    rc = H.qth_contract_sequence(qasm.encode(), measure.encode(), flat, len(pairs), v, ctypes.byref(flops), ctypes.byref(nodes), ctypes.byref(secs))
    if rc != 0:
        _raise(H, rc)
    return complex(v[0], v[1]), flops.value, nodes.value, secs.value


def export_plan_linegraph(qasm, measure, ordering, reduce=True):
    """-> (input_ranks, steps, inputs, flops): host bookkeeping only (no device), ready for Engine.plan(...)."""
    H = host_library()
    p = _QthPlan()
    rc = H.qth_export_plan_linegraph(qasm.encode(), measure.encode(), ordering.encode(), 1 if reduce else 0, ctypes.byref(p))
    if rc != 0:
        _raise(H, rc)
    ranks = [p.inputRanks[i] for i in range(p.nInputs)]
    inputs = []
    for i, r in enumerate(ranks):
        off = p.inputOffsets[i]
        arr = np.ctypeslib.as_array(p.inputData, shape=(2 * (off + 4 ** r),))[2 * off:]
        inputs.append(arr.copy().view(np.complex128))
    steps = []
    for i in range(p.nSteps):
        s = p.steps[i]
        steps.append((s.a, s.b, [s.pos_a[j] for j in range(s.k)], [s.pos_b[j] for j in range(s.k)]))
    return ranks, steps, inputs, p.flops


# ---- index slicing: the C++ host's planner and the sliced-amplitude executor (host/Slicing.h) ---------------------------

def _plan_steps(steps):
    arr = (PlanStep * max(len(steps), 1))()
    for i, (a, b, pa, pb) in enumerate(steps):
        arr[i].a, arr[i].b, arr[i].k = a, b, len(pa)
        for j, (x, y) in enumerate(zip(pa, pb)):
            arr[i].pos_a[j], arr[i].pos_b[j] = x, y
    return arr


def slice_plan(input_ranks, steps, n_wires):
    """host only: the C++ planner's choice of wires and its hoisted sliced plan.  Returns a dict with wires as
    (input tensor, leg) labels, ranks / steps of one slice, n_invariant, cuts {tensor: [(leg, wire), ...]}, units."""
    H = host_library()
    out = _QthSlicedPlan()
    ranks = (ctypes.c_int * max(len(input_ranks), 1))(*input_ranks)
    rc = H.qth_slice_plan(len(input_ranks), ranks, len(steps), _plan_steps(steps), n_wires, ctypes.byref(out))
    if rc != 0:
        _raise(H, rc)
    label = lambda w: (w // 32, w % 32)
    cuts = {}
    for i in range(out.nCuts):
        cuts.setdefault(out.cuts[3 * i], []).append((out.cuts[3 * i + 1], label(out.cuts[3 * i + 2])))
    st = []
    for i in range(out.nSteps):
        s = out.steps[i]
        st.append((s.a, s.b, [s.pos_a[j] for j in range(s.k)], [s.pos_b[j] for j in range(s.k)]))
    return {"wires": [label(out.wires[i]) for i in range(out.nWires)], "ranks": [out.inputRanks[i] for i in range(out.nInputs)],
            "steps": st, "n_invariant": out.nInvariant, "cuts": cuts, "peak_rank": out.peakRank,
            "units_per_slice": out.unitsPerSlice, "units_invariant": out.unitsInvariant}


def slice_tensor(full, rank, legs, digits):
    H = host_library()
    full = np.ascontiguousarray(full, dtype=np.complex128).ravel()
    out = np.empty(4 ** (rank - len(legs)), dtype=np.complex128)
    la, da = (ctypes.c_int * max(len(legs), 1))(*legs), (ctypes.c_int * max(len(legs), 1))(*digits)
    if H.qth_slice_tensor(full.ctypes.data, rank, la, da, len(legs), out.ctypes.data) != 0:
        _raise(H, 1)
    return out


class SlicedNetwork:
    """A line-graph network cut into 4^s slices dealt round-robin over the ranks of the job (SlicedContraction in
    host/Slicing.h): plan once on the host, then per amplitude stage -> begin -> end with no host synchronisation inside,
    slot scalars summed on the device and ONE in-stream ncclAllReduce.  ordering "" = in-process min-fill.
    With world > 1 the NCCL communicator must already be initialised on host_api.engine()."""

    def __init__(self, qasm, measure, ordering, reduce=True, slice_wires=-1, lanes=2, rank=0, world=1):
        self.H = host_library()
        self.h = self.H.qth_sliced_create(qasm.encode(), measure.encode(), (ordering or "").encode(), 1 if reduce else 0, slice_wires, lanes, rank, world)
        if not self.h:
            _raise(self.H, 1)
        info = (ctypes.c_longlong * 8)()
        ups, uinv, fl = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        self.H.qth_sliced_info(self.h, info, ctypes.byref(ups), ctypes.byref(uinv), ctypes.byref(fl))
        (self.slices, self.owned, self.invariant_steps, self.steps, self.peak_rank, self.cut_wires, self.launches_per_slice,
         self.launches_prefix) = (int(x) for x in info)
        self.units_per_slice, self.units_invariant, self.units_unsliced = ups.value, uinv.value, fl.value
        self.units_total = self.units_invariant + (self.units_per_slice - self.units_invariant) * self.slices

    def stage(self, bank=0):
        if self.H.qth_sliced_stage(self.h, bank) != 0:
            _raise(self.H, 1)

    def begin(self, bank=0):
        t = self.H.qth_sliced_begin(self.h, bank)
        if t < 0:
            _raise(self.H, 1)
        return t

    def end(self, ticket):
        v = (ctypes.c_double * 2)()
        if self.H.qth_sliced_end(self.h, ticket, v) != 0:
            _raise(self.H, 1)
        return complex(v[0], v[1])

    def amplitude(self):
        self.stage(0)
        return self.end(self.begin(0))

    def close(self):
        if self.h:
            self.H.qth_sliced_destroy(self.h)
            self.h = None
