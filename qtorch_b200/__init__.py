"""qtorch_b200 -- B200-native contraction engine behind qTorch's Network/Node API.

This Python package is only the thin ctypes face of ``libqtorch_b200.so`` (C ABI: include/qtorch_b200.h)
used by the tests and bench.py; the product is the shared library plus the C++14 host mirror in
``qtorch_b200/host`` (drop-in for /root/reference/src/*.h).  There is no CPU fallback: constructing an
``Engine`` without a B200 raises ``DeviceUnavailable``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqtorch_b200.so")
HARNESS_PATH = os.path.join(_HERE, "bin", "qtb_harness")
CLI_PATH = os.path.join(_HERE, "bin", "qtorch")
QTB_MAX_RANK = 16
QTB_UNIQUE_ID_BYTES = 128

# every symbol include/qtorch_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "qtb_abi_version", "qtb_status_string", "qtb_last_error", "qtb_device_count",
    "qtb_ctx_create", "qtb_ctx_destroy", "qtb_ctx_sync", "qtb_ctx_flush", "qtb_ctx_stream",
    "qtb_tensor_alloc", "qtb_tensor_free", "qtb_tensor_rank", "qtb_tensor_device_ptr",
    "qtb_tensor_upload", "qtb_tensor_download", "qtb_read_scalar", "qtb_read_scalar_begin", "qtb_read_scalar_end", "qtb_contract",
    "qtb_plan_create", "qtb_plan_destroy", "qtb_plan_run_host", "qtb_plan_upload_inputs",
    "qtb_plan_run_device", "qtb_plan_read_output", "qtb_plan_stage_inputs", "qtb_plan_run_device_slot", "qtb_plan_create_sliced", "qtb_plan_run_slots", "qtb_plan_prefix_units", "qtb_plans_run_batched", "qtb_plan_output_rank", "qtb_plan_units", "qtb_plan_launches",
    "qtb_sliced_create", "qtb_sliced_destroy", "qtb_sliced_stage", "qtb_sliced_begin", "qtb_sliced_lanes", "qtb_sliced_units", "qtb_sliced_prefix_units", "qtb_sliced_launches",
    "qtb_batch_create", "qtb_batch_destroy", "qtb_batch_set_inputs", "qtb_batch_bind", "qtb_batch_begin", "qtb_batch_end", "qtb_batch_run", "qtb_batch_launches", "qtb_batch_units",
    "qtb_comm_unique_id", "qtb_comm_init", "qtb_comm_destroy", "qtb_allreduce_sum", "qtb_allreduce_sum_device",
    "qtb_ctx_stats", "qtb_ctx_reset_stats", "qtb_ctx_timer_start", "qtb_ctx_timer_stop", "qtb_ctx_set_micro_limit", "qtb_ctx_get_micro_limit", "qtb_ctx_trace_enable", "qtb_ctx_trace_read",
]


class DeviceUnavailable(RuntimeError):
    pass


class EngineError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("qtorch_b200 status %d: %s" % (status, msg))
        self.status = status


class PlanStep(ctypes.Structure):
    _fields_ = [("a", ctypes.c_int32), ("b", ctypes.c_int32), ("k", ctypes.c_int32),
                ("pos_a", ctypes.c_int8 * QTB_MAX_RANK), ("pos_b", ctypes.c_int8 * QTB_MAX_RANK)]


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_longlong) for n in ("launches", "steps", "micro_steps", "units", "bytes_h2d", "bytes_d2h",
                                                  "pool_bytes_reserved", "pool_bytes_peak_live", "tma_launches")]


class StepTrace(ctypes.Structure):
    _fields_ = [("rank_a", ctypes.c_int32), ("rank_b", ctypes.c_int32), ("k", ctypes.c_int32), ("kernel", ctypes.c_int32),
                ("ms", ctypes.c_float)]


_lib = None


def load_library():
    """dlopen libqtorch_b200.so (raises if it has not been built: the product never falls back)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DeviceUnavailable("libqtorch_b200.so is not built (run `python -m qtorch_b200.build`)")
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cpi = ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)
    L.qtb_status_string.restype = ctypes.c_char_p
    L.qtb_status_string.argtypes = [ci]
    L.qtb_last_error.restype = ctypes.c_char_p
    L.qtb_device_count.argtypes = [cpi]
    L.qtb_ctx_create.argtypes = [ci, ctypes.POINTER(vp)]
    L.qtb_ctx_destroy.argtypes = [vp]
    L.qtb_ctx_sync.argtypes = [vp]
    L.qtb_ctx_flush.argtypes = [vp]
    L.qtb_ctx_stream.restype = vp
    L.qtb_ctx_stream.argtypes = [vp]
    L.qtb_tensor_alloc.argtypes = [vp, ci, ctypes.POINTER(vp)]
    L.qtb_tensor_free.argtypes = [vp, vp]
    L.qtb_tensor_rank.argtypes = [vp]
    L.qtb_tensor_device_ptr.restype = vp
    L.qtb_tensor_device_ptr.argtypes = [vp]
    L.qtb_tensor_upload.argtypes = [vp, vp, vp]
    L.qtb_tensor_download.argtypes = [vp, vp, vp]
    L.qtb_read_scalar.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_double)]
    L.qtb_read_scalar_begin.argtypes = [vp, vp, ctypes.POINTER(vp)]
    L.qtb_read_scalar_end.argtypes = [vp, vp, ctypes.POINTER(ctypes.c_double)]
    L.qtb_contract.argtypes = [vp, vp, vp, ci, cpi, cpi, vp]
    L.qtb_plan_create.argtypes = [vp, ci, cpi, ci, ctypes.POINTER(PlanStep), ctypes.POINTER(vp)]
    L.qtb_plan_destroy.argtypes = [vp, vp]
    L.qtb_plan_run_host.argtypes = [vp, vp, ctypes.POINTER(vp), vp]
    L.qtb_plan_upload_inputs.argtypes = [vp, vp, ctypes.POINTER(vp)]
    L.qtb_plan_run_device.argtypes = [vp, vp]
    L.qtb_plan_read_output.argtypes = [vp, vp, vp]
    L.qtb_plan_stage_inputs.argtypes = [vp, vp, ci, ctypes.POINTER(vp)]
    L.qtb_plan_run_device_slot.argtypes = [vp, vp, ci]
    L.qtb_plan_create_sliced.argtypes = [vp, ci, cpi, ci, ctypes.POINTER(PlanStep), ci, ctypes.POINTER(vp)]
    L.qtb_plan_run_slots.argtypes = [vp, vp, cpi, ci, vp, vp]
    L.qtb_plan_prefix_units.restype = ctypes.c_longlong
    L.qtb_plan_prefix_units.argtypes = [vp]
    L.qtb_ctx_set_micro_limit.argtypes = [vp, ci]
    L.qtb_ctx_get_micro_limit.argtypes = [vp]
    L.qtb_ctx_timer_start.argtypes = [vp]
    L.qtb_ctx_timer_stop.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    L.qtb_plan_output_rank.argtypes = [vp]
    L.qtb_plan_units.restype = ctypes.c_longlong
    L.qtb_plan_units.argtypes = [vp]
    L.qtb_plan_launches.argtypes = [vp]
    L.qtb_comm_unique_id.argtypes = [ctypes.c_char_p]
    L.qtb_comm_init.argtypes = [vp, ci, ci, ctypes.c_char_p]
    L.qtb_comm_destroy.argtypes = [vp]
    L.qtb_allreduce_sum.argtypes = [vp, vp, ci]
    L.qtb_ctx_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    L.qtb_ctx_reset_stats.argtypes = [vp]
    L.qtb_ctx_trace_enable.argtypes = [vp, ci]
    L.qtb_ctx_trace_read.argtypes = [vp, ctypes.POINTER(StepTrace), ci, cpi]
    _lib = L
    return L


def _check(status):
    if status == 0:
        return
    L = load_library()
    msg = "%s (%s)" % (L.qtb_status_string(status).decode(), L.qtb_last_error().decode())
    if status == 1:
        raise DeviceUnavailable(msg)
    raise EngineError(status, msg)


def _iarr(v):
    return (ctypes.c_int * max(len(v), 1))(*v)


class Tensor:
    """Device tensor handle (rank r <-> 4^r complex128, leg 0 fastest)."""

    def __init__(self, engine, rank):
        self.engine, self.rank = engine, rank
        h = ctypes.c_void_p()
        _check(engine.lib.qtb_tensor_alloc(engine.ctx, rank, ctypes.byref(h)))
        self.handle = h

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=np.complex128).ravel()
        assert host.size == 4 ** self.rank
        _check(self.engine.lib.qtb_tensor_upload(self.engine.ctx, self.handle, host.ctypes.data))
        return self

    def download(self):
        out = np.empty(4 ** self.rank, dtype=np.complex128)
        _check(self.engine.lib.qtb_tensor_download(self.engine.ctx, self.handle, out.ctypes.data))
        return out

    def scalar(self):
        v = (ctypes.c_double * 2)()
        _check(self.engine.lib.qtb_read_scalar(self.engine.ctx, self.handle, v))
        return complex(v[0], v[1])

    def free(self):
        if self.handle:
            _check(self.engine.lib.qtb_tensor_free(self.engine.ctx, self.handle))
            self.handle = None


class Plan:
    """Compiled contraction plan.  ``steps`` = [(a, b, posA, posB), ...] with tensor ids in the reference's
    mCreatedFrom numbering (inputs 0..n-1, result of step i = n+i)."""

    def __init__(self, engine, input_ranks, steps, invariant_steps=0):
        self.engine = engine
        self.input_ranks = list(input_ranks)
        arr = (PlanStep * len(steps))()
        for i, (a, b, pa, pb) in enumerate(steps):
            arr[i].a, arr[i].b, arr[i].k = a, b, len(pa)
            for j, (x, y) in enumerate(zip(pa, pb)):
                arr[i].pos_a[j], arr[i].pos_b[j] = x, y
        h = ctypes.c_void_p()
        _check(engine.lib.qtb_plan_create_sliced(engine.ctx, len(input_ranks), _iarr(self.input_ranks), len(steps), arr,
                                                 int(invariant_steps), ctypes.byref(h)))
        self.handle = h
        self.n_steps = len(steps)

    def _ptrs(self, host_inputs):
        self._keep = [np.ascontiguousarray(x, dtype=np.complex128).ravel() for x in host_inputs]
        assert len(self._keep) == len(self.input_ranks)
        return (ctypes.c_void_p * max(len(self._keep), 1))(*[x.ctypes.data for x in self._keep])

    @property
    def output_rank(self):
        return self.engine.lib.qtb_plan_output_rank(self.handle)

    @property
    def units(self):
        return self.engine.lib.qtb_plan_units(self.handle)

    @property
    def launches(self):
        return self.engine.lib.qtb_plan_launches(self.handle)

    def run_host(self, host_inputs):
        out = np.empty(4 ** self.output_rank, dtype=np.complex128)
        _check(self.engine.lib.qtb_plan_run_host(self.engine.ctx, self.handle, self._ptrs(host_inputs), out.ctypes.data))
        return out

    def upload_inputs(self, host_inputs):
        _check(self.engine.lib.qtb_plan_upload_inputs(self.engine.ctx, self.handle, self._ptrs(host_inputs)))

    def run_device(self):
        _check(self.engine.lib.qtb_plan_run_device(self.engine.ctx, self.handle))

    def stage_inputs(self, slot, host_inputs):
        _check(self.engine.lib.qtb_plan_stage_inputs(self.engine.ctx, self.handle, slot, self._ptrs(host_inputs)))

    def run_device_slot(self, slot):
        _check(self.engine.lib.qtb_plan_run_device_slot(self.engine.ctx, self.handle, slot))

    def run_slots(self, slots, each=False):
        """Sum of the scalar outputs over the staged input slots: invariant prefix once, the rest per slot, one sync."""
        slots = list(slots)
        total = np.zeros(1, dtype=np.complex128)
        per = np.zeros(max(len(slots), 1), dtype=np.complex128)
        _check(self.engine.lib.qtb_plan_run_slots(self.engine.ctx, self.handle, _iarr(slots), len(slots), total.ctypes.data, per.ctypes.data))
        return (complex(total[0]), per[:len(slots)]) if each else complex(total[0])

    @property
    def prefix_units(self):
        return self.engine.lib.qtb_plan_prefix_units(self.handle)

    def read_output(self):
        out = np.empty(4 ** self.output_rank, dtype=np.complex128)
        _check(self.engine.lib.qtb_plan_read_output(self.engine.ctx, self.handle, out.ctypes.data))
        return out

    def destroy(self):
        if self.handle:
            _check(self.engine.lib.qtb_plan_destroy(self.engine.ctx, self.handle))
            self.handle = None


class Engine:
    """One qtb_ctx: device, stream, pooled tensor storage, deferred micro-steps."""

    def __init__(self, device=None, ctx=None):
        self.lib = load_library()
        if ctx is not None:                       # wrap an existing qtb_ctx (e.g. the host mirror's singleton)
            self.ctx, self.device, self._borrowed = ctypes.c_void_p(ctx), device, True
            return
        self._borrowed = False
        if device is None:
            device = int(os.environ.get("QTORCH_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        h = ctypes.c_void_p()
        _check(self.lib.qtb_ctx_create(device, ctypes.byref(h)))
        self.ctx, self.device = h, device

    def tensor(self, rank, host=None):
        t = Tensor(self, rank)
        if host is not None:
            t.upload(host)
        return t

    def contract(self, a, b, pos_a, pos_b, out=None):
        """C = A (x) B over the shared legs (reference Network::ContractIndices); asynchronous."""
        k = len(pos_a)
        c = out if out is not None else Tensor(self, a.rank + b.rank - 2 * k)
        _check(self.lib.qtb_contract(self.ctx, a.handle, b.handle, k, _iarr(pos_a), _iarr(pos_b), c.handle))
        return c

    def plan(self, input_ranks, steps, invariant_steps=0):
        return Plan(self, input_ranks, steps, invariant_steps)

    def sync(self):
        _check(self.lib.qtb_ctx_sync(self.ctx))

    def flush(self):
        _check(self.lib.qtb_ctx_flush(self.ctx))

    @property
    def stream(self):
        return self.lib.qtb_ctx_stream(self.ctx)

    def stats(self):
        s = Stats()
        _check(self.lib.qtb_ctx_stats(self.ctx, ctypes.byref(s)))
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    def reset_stats(self):
        _check(self.lib.qtb_ctx_reset_stats(self.ctx))

    def set_micro_limit(self, log4_units):
        _check(self.lib.qtb_ctx_set_micro_limit(self.ctx, log4_units))

    def timer_start(self):
        _check(self.lib.qtb_ctx_timer_start(self.ctx))

    def timer_stop(self):
        ms = ctypes.c_float()
        _check(self.lib.qtb_ctx_timer_stop(self.ctx, ctypes.byref(ms)))
        return ms.value

    def trace(self, on=True):
        _check(self.lib.qtb_ctx_trace_enable(self.ctx, 1 if on else 0))

    def read_trace(self, max_entries=65536):
        arr = (StepTrace * max_entries)()
        n = ctypes.c_int()
        _check(self.lib.qtb_ctx_trace_read(self.ctx, arr, max_entries, ctypes.byref(n)))
        return [dict(rank_a=arr[i].rank_a, rank_b=arr[i].rank_b, k=arr[i].k, kernel=arr[i].kernel, ms=arr[i].ms) for i in range(n.value)]

    # ---- multi-GPU scalar reduction (NCCL) ----
    def comm_unique_id(self):
        buf = ctypes.create_string_buffer(QTB_UNIQUE_ID_BYTES)
        _check(self.lib.qtb_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, n_ranks, rank, uid):
        _check(self.lib.qtb_comm_init(self.ctx, n_ranks, rank, uid))

    def allreduce_sum(self, values):
        v = np.ascontiguousarray(values, dtype=np.complex128).ravel().copy()
        _check(self.lib.qtb_allreduce_sum(self.ctx, v.ctypes.data, v.size))
        return v

    def close(self):
        if self.ctx and not self._borrowed:
            self.lib.qtb_ctx_destroy(self.ctx)
        self.ctx = None


# ---- host-mirror drivers (C++ binaries) -----------------------------------------------------------------------

def run_harness(args, plan_only=False, cwd=None, timeout=None, extra_env=None):
    """Run qtorch_b200/bin/qtb_harness and parse its ``@@`` lines (same format as oracle/ref_harness)."""
    env = dict(os.environ)
    env["QTORCH_QUIET"] = "1"
    if plan_only:
        env["QTORCH_PLAN_ONLY"] = "1"
    if extra_env:
        env.update(extra_env)
    p = subprocess.run([HARNESS_PATH] + [str(a) for a in args], cwd=cwd, capture_output=True, text=True, timeout=timeout, env=env)
    out = {"_rc": p.returncode, "pstep": [], "_stdout": p.stdout, "_stderr": p.stderr}
    for line in p.stdout.splitlines():
        if not line.startswith("@@"):
            continue
        parts = line[2:].split()
        if not parts:
            continue
        if parts[0] == "pstep":
            out["pstep"].append([int(x) for x in parts[1:]])
        elif parts[0] == "exception":
            out["exception"] = " ".join(parts[1:])
        else:
            out[parts[0]] = parts[1:]
    return out


def plan_from_harness(out):
    """(input_ranks, steps) for Engine.plan from a ``+steps`` harness run."""
    ranks = [int(x) for x in out["inputs"][1:]]
    n = int(out["inputs"][0])
    assert n == len(ranks)
    steps = []
    for i, ps in enumerate(out["pstep"]):
        a, b, c, ra, rb, rc, k = ps[:7]
        assert c == n + i, "plan ids are not sequential"
        steps.append((a, b, ps[7:7 + k], ps[7 + k:7 + 2 * k]))
    return ranks, steps
