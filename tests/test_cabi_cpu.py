"""CPU-side checks of the drop-in boundary: libthzgpu.so loads, exports every symbol that
include/thzgpu.h declares, and refuses to compute without a GPU (no silent fallback)."""
import ctypes
import os

import pytest

from helpers import pkg


def test_library_exports_every_declared_symbol():
    m = pkg()
    L = m.load_library()
    assert len(m.DECLARED_SYMBOLS) >= 20
    missing = [s for s in m.DECLARED_SYMBOLS if not hasattr(L, s)]
    assert not missing, f"declared in include/thzgpu.h but not exported: {missing}"
    unbound = [s for s in m.DECLARED_SYMBOLS if s not in L._signatures]
    assert not unbound, f"no ctypes signature for: {unbound}"


def test_no_cpu_fallback_without_device():
    m = pkg()
    L = m.load_library()
    if L.thz_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(m.ThzError) as ei:
        m.Context(0)
    assert "no CUDA device" in str(ei.value)
    # compute entry points reject a null context instead of doing anything
    assert L.thz_trace_fused_dev(None, None, None, None, 0) != 0


def test_product_package_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pk = os.path.join(root, "thz-image-explorer_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "thz_oracle" not in txt, f
