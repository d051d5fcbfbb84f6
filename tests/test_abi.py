"""The C-ABI boundary: libqtorch_b200.so loads without a GPU, exports every symbol include/qtorch_b200.h
declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
import qtorch_b200 as qt


def _header_functions():
    text = open(os.path.join(ROOT, "include", "qtorch_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qtb_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree(built):
    assert _header_functions() == sorted(qt.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(qt.LIB_PATH)
    for name in _header_functions():
        assert hasattr(L, name), name
    assert L.qtb_abi_version() == 1


def test_struct_layouts_match_header(built):
    assert ctypes.sizeof(qt.PlanStep) == 12 + 2 * qt.QTB_MAX_RANK
    assert ctypes.sizeof(qt.Stats) == 9 * 8
    assert ctypes.sizeof(qt.StepTrace) == 20


def test_no_cpu_fallback(built):
    """On a box without a GPU the product must fail loudly; on a GPU box this test is a no-op."""
    L = qt.load_library()
    n = ctypes.c_int()
    if L.qtb_device_count(ctypes.byref(n)) == 0 and n.value > 0:
        pytest.skip("GPU present")
    with pytest.raises(qt.DeviceUnavailable):
        qt.Engine(0)
    # the host mirror refuses too: the harness exits through DeviceUnavailable, never computes on the CPU
    from conftest import GOLDEN
    out = qt.run_harness(["lg", os.path.join(GOLDEN, "Samples/bell_pair.qasm"), os.path.join(GOLDEN, "measure/bell_00.txt"),
                          os.path.join(GOLDEN, "orderings/bell_00.qbb.out"), 1])
    assert "device engine unavailable" in out.get("exception", "")


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under qtorch_b200/ may reference it"""
    for base, _, files in os.walk(os.path.join(ROOT, "qtorch_b200")):
        for f in files:
            if f.endswith((".py", ".h", ".hpp", ".cu", ".cuh", ".cpp", ".inl")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "contract_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_micro_step_lane_layout_invariants(built):
    """host logic of the grouped micro-step executor (engine.cu micro_item_layout): for every step shape and every target chain
    length the layout must tile the outputs exactly (items x passes x outputs-per-pass = 4^rC), use whole warps, never put more
    lanes on an output than there are summed terms unless the result itself is smaller than a warp, and return the chain length
    the kernel will really run."""
    L = qt.load_library()
    L.qtb_debug_micro_layout.restype = ctypes.c_int
    L.qtb_debug_micro_layout.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    for rC in range(0, 8):
        for k in range(0, 9):
            NC, K = 4 ** rC, 4 ** k
            prev = None
            for target in (1, 4, 16, 64, 256, 4096, 10 ** 6):
                lg, lp = ctypes.c_int(), ctypes.c_int()
                serial = L.qtb_debug_micro_layout(rC, k, target, ctypes.byref(lg), ctypes.byref(lp))
                G, P, passes = 1 << lg.value, 32 >> lg.value, 1 << lp.value
                assert 0 <= lg.value <= 5 and 0 <= lp.value <= 2
                assert P <= NC or NC < 32 and P == NC, (rC, k, target, lg.value)
                assert P * passes <= NC and NC % (P * passes) == 0
                assert G <= max(K, 32 // min(NC, 32)), (rC, k, target, G)          # more lanes than terms only to fill a warp
                assert serial == passes * max(1, K >> lg.value)
                if prev is not None:
                    assert serial >= prev                                          # a looser target never shortens the chain
                prev = serial
            assert prev == min(4, max(1, NC // min(NC, 32))) * max(1, K >> max(0, 5 - min(5, 2 * rC)))   # no pressure: 4 passes (or all outputs), G = 32 / NC at most
