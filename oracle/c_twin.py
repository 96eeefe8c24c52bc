"""ctypes loader of the oracle's C / pthreads twin (oracle/thz_oracle_c.c).  Test infrastructure and
CPU baseline only -- never imported by the product package."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_twin", "libthzoracle.so")
        if not os.path.exists(path):   # build() compiles it; do it on demand for a bare `pytest`
            import subprocess
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        L = C.CDLL(path)
        fp = C.c_void_p
        L.thzc_default_chain.restype = C.c_int
        L.thzc_default_chain.argtypes = [fp, C.c_int, C.c_int, C.c_int, fp, fp, fp, fp, fp, fp, fp, fp, fp, fp, C.c_int]
        L.thzc_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def default_chain(cube, tilt, gate_before, window, band, gate_after, threads=0, want_spectra=False):
    """Slots 1..7 of the default chain -> (data7, img[, fft5, amp5, phase4])."""
    cube = np.ascontiguousarray(cube, np.float32)
    rows, cols, n = cube.shape
    F = n // 2 + 1
    out = np.empty_like(cube)
    img = np.empty((rows, cols), np.float32)

    def vp(a):
        return None if a is None else np.ascontiguousarray(a, np.float32).ctypes.data

    keep = [np.ascontiguousarray(a, np.float32) if a is not None else None
            for a in (tilt, gate_before, window, band, gate_after)]
    fft5 = np.empty((rows, cols, F), np.complex64) if want_spectra else None
    amp5 = np.empty((rows, cols, F), np.float32) if want_spectra else None
    ph4 = np.empty((rows, cols, F), np.float32) if want_spectra else None
    rc = lib().thzc_default_chain(cube.ctypes.data, rows, cols, n, *[None if k is None else k.ctypes.data for k in keep],
                                  out.ctypes.data, img.ctypes.data,
                                  None if fft5 is None else fft5.ctypes.data,
                                  None if amp5 is None else amp5.ctypes.data,
                                  None if ph4 is None else ph4.ctypes.data, int(threads))
    assert rc == 0
    return (out, img, fft5, amp5, ph4) if want_spectra else (out, img)
