"""ctypes/numpy front-end to the CPU oracle (TEST INFRASTRUCTURE ONLY).

* ``contract``          -- C restatement of Network::ContractIndices
                           (/root/reference/src/Network.h:876-971) via oracle/_build/libcontract_oracle.so
* ``contract_numpy``    -- independent numpy restatement (einsum on Fortran-ordered views), used to
                           cross-check the C restatement on small cases
* ``ref_harness``       -- run the UNMODIFIED reference (oracle/_ref/ref_harness) and parse its ``@@`` lines
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile the C restatement (and, where /root/reference exists, oracle/_ref)."""
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)
    if os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libcontract_oracle.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.qto_contract.restype = ctypes.c_int
        L.qto_contract.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                   ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]
        L.qto_step_units.restype = ctypes.c_longlong
        L.qto_step_units.argtypes = [ctypes.c_int] * 3
        L.qto_final_value_rule.restype = ctypes.c_int
        L.qto_final_value_rule.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int]
        _LIB = L
    return _LIB


def _iarr(v):
    return (ctypes.c_int * max(len(v), 1))(*v)


def contract(A, rA, B, rB, posA, posB):
    """C[c] = sum_s A*B with the reference's leg conventions; returns a complex128 vector of 4^rC."""
    A = np.ascontiguousarray(A, dtype=np.complex128).ravel()
    B = np.ascontiguousarray(B, dtype=np.complex128).ravel()
    k = len(posA)
    assert A.size == 4 ** rA and B.size == 4 ** rB and len(posB) == k
    C = np.empty(4 ** (rA + rB - 2 * k), dtype=np.complex128)
    rc = lib().qto_contract(A.ctypes.data, rA, B.ctypes.data, rB, k, _iarr(posA), _iarr(posB), C.ctypes.data)
    if rc != 0:
        raise ValueError("bad leg maps")
    return C


def step_units(rA, rB, k):
    return lib().qto_step_units(rA, rB, k)


def contract_numpy(A, rA, B, rB, posA, posB):
    """Independent restatement: tensordot over Fortran-ordered (leg 0 fastest) views."""
    At = np.asarray(A, dtype=np.complex128).reshape((4,) * rA, order="F") if rA else np.asarray(A).reshape(())
    Bt = np.asarray(B, dtype=np.complex128).reshape((4,) * rB, order="F") if rB else np.asarray(B).reshape(())
    Ct = np.tensordot(At, Bt, axes=(list(posA), list(posB)))  # free A legs (A order) then free B legs (B order)
    return np.asarray(Ct).reshape(-1, order="F")


def ref_harness_path(shipped=False):
    return os.path.join(_HERE, "_ref", "ref_harness_shipped" if shipped else "ref_harness")


def ref_available():
    return os.path.exists(ref_harness_path())


def ref_harness(args, cwd=None, shipped=False, timeout=None, env=None):
    """Run the unmodified reference through oracle/_ref/ref_harness; return {tag: [fields...]} of @@ lines
    (``step`` lines accumulate into a list)."""
    p = subprocess.run([ref_harness_path(shipped)] + [str(a) for a in args], cwd=cwd, capture_output=True,
                       text=True, timeout=timeout, env=env)
    out = {"_rc": p.returncode, "step": []}
    for line in p.stdout.splitlines():
        if not line.startswith("@@"):
            continue
        parts = line[2:].split()
        if not parts:
            continue
        if parts[0] in ("step", "term"):
            out.setdefault(parts[0], []).append(parts[1:])
        else:
            out[parts[0]] = parts[1:]
    if p.returncode not in (0, 1):
        raise RuntimeError("ref_harness failed: " + p.stderr[-500:])
    return out


def ref_step(A, rA, B, rB, posA, posB, tmpdir):
    """One step through the reference's own Network::ContractNodes."""
    fa, fb, fc = (os.path.join(tmpdir, n) for n in ("A.bin", "B.bin", "C.bin"))
    np.ascontiguousarray(A, dtype=np.complex128).tofile(fa)
    np.ascontiguousarray(B, dtype=np.complex128).tofile(fb)
    k = len(posA)
    out = ref_harness(["step", rA, rB, k] + list(posA) + list(posB) + [fa, fb, fc])
    if "rejected" in out:
        return None, out
    return np.fromfile(fc, dtype=np.complex128), out
