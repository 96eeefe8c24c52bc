"""thz-image-explorer_b200 -- B200 (sm_100a) implementation of the thz-image-explorer
filter-chain hot path.

The product is ``libthzgpu.so`` (hand-written CUDA behind the C ABI of
``include/thzgpu.h``).  This package is the thin Python host side used by the tests and
``bench.py``: a ``ctypes`` binding (``lib``) and ``Context``, a numpy-in / numpy-out
mirror of the reference's stage functions (``math_tools::fft`` / ``ifft``, the band-pass
and time-gate filters, ``Deconvolution``).  There is no CPU fallback anywhere in this
package: if the shared library or a B200 is missing, calls raise ``ThzError``.

Import it with ``importlib.import_module("thz-image-explorer_b200")`` (the directory
name is not a Python identifier).
"""
from __future__ import annotations

from .binding import (Chain, Context, DeviceBuffer, Group, Slab, ThzError, lib, library_path, load_library,  # noqa: F401
                      DECLARED_SYMBOLS)
from . import host  # noqa: F401,E402
from . import sharding  # noqa: F401,E402
