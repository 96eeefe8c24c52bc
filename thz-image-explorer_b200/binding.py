"""ctypes binding of libthzgpu.so (include/thzgpu.h) and a numpy-facing ``Context``.

No computation happens in this file: every method forwards to one C-ABI entry point.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
HEADER = os.path.join(_ROOT, "include", "thzgpu.h")

THZ_OK, THZ_ABORTED = 0, 1


class ThzError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libthzgpu error {code}: {msg}")
        self.code = code


def library_path() -> str:
    return os.path.join(_HERE, "libthzgpu.so")


def declared_symbols(header: str = HEADER):
    """Names of all functions declared in include/thzgpu.h."""
    txt = open(header).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(thz_[a-z0-9_]+)\s*\(", txt)) - {"thz_progress_fn"})


def _as_np(ptr, shape, dtype=np.float32):
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return np.zeros(shape, dtype)
    buf = (C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()


DECLARED_SYMBOLS = declared_symbols() if os.path.exists(HEADER) else []

THZ_MAX_PSF = 255
THZ_FIR_TAPS = 499
THZ_MAX_BANDS = 32
SKIP_REASONS = {2: "no dx/dy", 3: "no psf", 4: "image too small", 5: "psf too large"}


class SplineC(C.Structure):
    _fields_ = [("n", C.c_int), ("knots", C.c_void_p), ("values", C.c_void_p), ("coeff_a", C.c_void_p),
                ("coeff_b", C.c_void_p), ("coeff_c", C.c_void_p), ("coeff_d", C.c_void_p)]


class HybridFitC(C.Structure):
    _fields_ = [("base_a", C.c_float), ("base_b", C.c_float), ("correction", SplineC)]


class PsfC(C.Structure):
    _fields_ = [("wx_fit", HybridFitC), ("wy_fit", HybridFitC), ("x0_spline", SplineC), ("y0_spline", SplineC)]


class DeconvParamsC(C.Structure):
    _fields_ = [("n_iterations", C.c_int), ("n_filters", C.c_int), ("start_freq", C.c_float),
                ("end_freq", C.c_float), ("win_width", C.c_float)]


class BandPlanC(C.Structure):
    _fields_ = [("center_freq", C.c_float), ("wx", C.c_float), ("wy", C.c_float), ("x0", C.c_float),
                ("y0", C.c_float), ("kx", C.c_int), ("ky", C.c_int), ("n_iter", C.c_int), ("direct", C.c_int),
                ("psf_x", C.c_float * THZ_MAX_PSF), ("psf_y", C.c_float * THZ_MAX_PSF),
                ("fir", C.c_float * THZ_FIR_TAPS)]

    def psf_x_np(self):
        return np.ctypeslib.as_array(self.psf_x)[: self.kx].copy()

    def psf_y_np(self):
        return np.ctypeslib.as_array(self.psf_y)[: self.ky].copy()

    def fir_np(self):
        return np.ctypeslib.as_array(self.fir).copy()


PROGRESS_FN = C.CFUNCTYPE(None, C.c_float, C.c_void_p)

_lib = None


def load_library():
    """dlopen libthzgpu.so; raises ThzError when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise ThzError(-2, f"{path} not found: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(path)
    vp, i32, i64, u64, f32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_size_t
    fp = C.c_void_p  # float* (host or device), passed as integer addresses
    sig = {
        "thz_device_count": (i32, []),
        "thz_ctx_create": (i32, [i32, C.POINTER(vp)]),
        "thz_ctx_destroy": (None, [vp]),
        "thz_last_error": (C.c_char_p, [vp]),
        "thz_ctx_device": (i32, [vp]),
        "thz_ctx_sm_count": (i32, [vp]),
        "thz_ctx_stream": (vp, [vp]),
        "thz_sync": (i32, [vp]),
        "thz_launch_count": (i64, [vp]),
        "thz_dev_alloc": (i32, [vp, sz, C.POINTER(vp)]),
        "thz_dev_free": (i32, [vp, vp]),
        "thz_dev_memset": (i32, [vp, vp, i32, sz]),
        "thz_copy_h2d": (i32, [vp, vp, vp, sz]),
        "thz_copy_d2h": (i32, [vp, vp, vp, sz]),
        "thz_host_alloc": (i32, [sz, C.POINTER(vp)]),
        "thz_host_free": (i32, [vp]),
        "thz_generate_cube": (i32, [vp, fp, i32, i32, i32, i32, i32, u64, f32, f32, f32]),
        "thz_frequency_axis": (i32, [fp, i32, fp]),
        "thz_adapted_blackman": (i32, [fp, i32, f32, f32, fp]),
        "thz_window_multiplier": (i32, [i32, fp, i32, f32, f32, fp]),
        "thz_time_gate_multiplier": (i32, [fp, i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, fp,
                                           C.POINTER(i32), C.POINTER(i32)]),
        "thz_band_pass_multiplier": (i32, [fp, i32, C.c_double, C.c_double, C.c_double, fp, C.POINTER(i32),
                                           C.POINTER(i32)]),
        "thz_plan_trace": (i32, [vp, i32, fp, fp, fp]),
        "thz_trace_fused_dev": (i32, [vp, fp, fp, fp, i64]),
        "thz_trace_forward_dev": (i32, [vp, fp, fp, fp, fp, fp, i64]),
        "thz_band_apply_dev": (i32, [vp, fp, fp, i64]),
        "thz_trace_inverse_dev": (i32, [vp, fp, i32, i32, fp, fp, i64]),
        "thz_spectral_means": (i32, [vp, fp, fp, fp, i64, fp, fp, fp]),
        "thz_fir_bank": (i32, [i32, C.c_double, C.c_double, C.c_double, f32, f32, fp, fp]),
        "thz_hybrid_eval": (f32, [C.POINTER(HybridFitC), f32]),
        "thz_spline_eval_const_extrap": (f32, [C.POINTER(SplineC), f32]),
        "thz_deconv_plan_bands": (i32, [C.POINTER(PsfC), C.POINTER(DeconvParamsC), fp, i32, i32, i32, i32, f32, f32,
                                        C.POINTER(BandPlanC)]),
        "thz_deconv_energies_dev": (i32, [vp, fp, i64, i32, C.POINTER(BandPlanC), i32, fp]),
        "thz_rl_separable_dev": (i32, [vp, fp, i32, i32, fp, i32, fp, i32, i32, i32, fp, fp, vp, vp, vp, f32, f32]),
        "thz_rl_dense_dev": (i32, [vp, fp, i32, i32, fp, i32, i32, i32, i32, fp, fp, vp]),
        "thz_conv2d_separable_dev": (i32, [vp, fp, i32, i32, fp, i32, fp, i32, i32, fp]),
        "thz_conv2d_dense_dev": (i32, [vp, fp, i32, i32, fp, i32, i32, i32, fp]),
        "thz_deconv_apply_dev": (i32, [vp, fp, fp, i64, i32, C.POINTER(BandPlanC), i32, fp, fp]),
        "thz_deconvolution_dev": (i32, [vp, fp, i32, i32, i32, C.POINTER(BandPlanC), i32, fp, fp, vp, vp, vp]),
        "thz_deconv_stage_ms": (i32, [vp, fp]),
        "thz_deconv_kernel_ms": (i32, [vp, fp]),
        "thz_chain_host": (i32, [vp, fp, i32, i32, i32, C.POINTER(BandPlanC), i32, fp, fp, vp, vp, vp]),
        "thz_chain_dev": (i32, [vp, fp, i32, i32, i32, C.POINTER(BandPlanC), i32, fp, fp, vp, vp, vp]),
        "thz_chain_energies_dev": (i32, [vp, fp, fp, fp, i64, i32, C.POINTER(BandPlanC), i32, fp]),
        "thz_chain_kernel_ms": (i32, [vp, fp]),
        "thz_chain_begin_dev": (i32, [vp, fp, fp, fp, i64, i32, C.POINTER(BandPlanC), i32, fp]),
        "thz_chain_end_dev": (i32, [vp, fp, fp, i64, i32, C.POINTER(BandPlanC), i32, fp, fp]),
        "thz_deconvolution_host": (i32, [vp, fp, i32, i32, i32, C.POINTER(BandPlanC), i32, fp, fp, vp, vp, vp]),
        "thz_scale_blocks_dev": (i32, [vp, fp, i32, i32, i32, i32, fp]),
        "thz_scale_blocks_host": (i32, [vp, fp, i32, i32, i32, i32, fp]),
        "thz_bias_subtract_dev": (i32, [vp, fp, i32, fp, fp, i64]),
        "thz_roi_average_dev": (i32, [vp, fp, i32, i32, i32, fp, fp, i32, i32, fp]),
        "thz_optical_properties": (i32, [fp, fp, fp, fp, fp, i32, f32, fp, fp, fp]),
        "thz_tilt_plan": (i32, [fp, i32, i32, i32, f32, f32, C.c_double, C.c_double, C.POINTER(i32), fp, fp]),
        "thz_tilt_shift_host": (i32, [vp, fp, fp, fp, i32, i32, fp, i64]),
        "thz_reference_pulse": (i32, [vp, fp, i32, fp, fp, i32, i32, f32, f32, fp, fp, fp]),
        "thz_voxel_opacity_dev": (i32, [vp, fp, i32, i64, f32, f32, f32, i32, i64, fp, fp]),
        "thz_time_multiply_dev": (i32, [vp, fp, fp, i32, fp, i64]),
        "thz_time_multiply_host": (i32, [vp, fp, fp, i32, fp, i64]),
        "thz_band_apply_host": (i32, [vp, fp, fp, i64]),
        "thz_spectral_means_host": (i32, [vp, fp, fp, fp, i64, fp, fp, fp]),
        "thz_intensity_host": (i32, [vp, fp, i32, fp, i64]),
        "thz_chain_create": (i32, [vp, C.POINTER(vp)]),
        "thz_chain_destroy": (None, [vp]),
        "thz_chain_length": (i32, [vp]),
        "thz_chain_stage_name": (C.c_char_p, [vp, i32]),
        "thz_chain_set_config": (i32, [vp, f32, f32, i32, i32]),
        "thz_chain_set_psf": (i32, [vp, C.POINTER(PsfC)]),
        "thz_chain_set_param": (i32, [vp, C.c_char_p, C.c_char_p, C.c_double]),
        "thz_chain_get_param": (i32, [vp, C.c_char_p, C.c_char_p, C.POINTER(C.c_double)]),
        "thz_chain_set_active": (i32, [vp, C.c_char_p, i32]),
        "thz_chain_open": (i32, [vp, fp, i32, fp, i32, i32, i32, f32, f32]),
        "thz_chain_run": (i32, [vp, i32, i32]),
        "thz_chain_run_fused": (i32, [vp, i32]),
        "thz_chain_abort": (None, [vp, i32]),
        "thz_chain_slot": (i32, [vp, i32] + [C.POINTER(vp)] * 8 + [C.POINTER(i32), C.POINTER(i32)]),
        "thz_chain_fused_result": (i32, [vp, C.POINTER(vp), C.POINTER(vp)]),
        "thz_chain_filter_ms": (C.c_double, [vp, C.c_char_p]),
        "thz_slab_create": (i32, [vp, i32, i32, C.POINTER(vp)]),
        "thz_slab_destroy": (None, [vp]),
        "thz_slab_plan": (i32, [vp, C.POINTER(i32), i32, C.POINTER(BandPlanC), i32, C.POINTER(i32)]),
        "thz_slab_export": (i32, [vp, vp]),
        "thz_slab_connect_ipc": (i32, [vp, vp]),
        "thz_slab_connect_local": (i32, [vp, vp, vp]),
        "thz_slab_set_stream": (i32, [vp, vp]),
        "thz_slab_rl": (i32, [vp, fp, i64, fp, fp]),
        "thz_slab_rl_serial": (i32, [C.POINTER(vp), i32, C.POINTER(vp), C.POINTER(i64), C.POINTER(vp), C.POINTER(vp)]),
        "thz_slab_status": (i32, [vp]),
        "thz_plan_reference": (i32, [vp, fp, fp, i32]),
        "thz_trace_forward_normalised_dev": (i32, [vp, fp, fp, fp, fp, fp, i64]),
        "thz_spectral_slice_dev": (i32, [vp, fp, i32, i32, fp, i64]),
        "thz_pixel_handoff_dev": (i32, [vp, fp, fp, i64, i64, fp, fp, fp, fp, fp]),
        "thz_mean_trace_dev": (i32, [vp, fp, i32, i64, fp]),
        "thz_mean_spectra_dev": (i32, [vp, fp, i64, fp, fp, fp]),
        "thz_kernel_timing_begin": (i32, [vp]),
        "thz_kernel_timing_end": (i32, [vp, fp]),
        "thz_chain_host_begin": (i32, [vp, fp, i32, i32, i32, C.POINTER(BandPlanC), i32, fp, C.POINTER(vp), C.POINTER(vp)]),
        "thz_chain_host_end": (i32, [vp, i32, i32, i32, C.POINTER(BandPlanC), i32, fp, fp]),
        "thz_fp32_rate": (i32, [vp, i32, C.POINTER(C.c_double)]),
        "thz_group_create": (i32, [C.POINTER(i32), i32, C.POINTER(vp)]),
        "thz_group_destroy": (None, [vp]),
        "thz_group_size": (i32, [vp]),
        "thz_group_ctx": (vp, [vp, i32]),
        "thz_group_last_error": (C.c_char_p, [vp]),
        "thz_group_row_bounds": (i32, [vp, i32, C.POINTER(i32)]),
        "thz_group_rl_host": (i32, [vp, fp, i32, i32, C.POINTER(BandPlanC), i32, fp]),
        "thz_group_chain_host": (i32, [vp, fp, i32, i32, i32, fp, fp, fp, C.POINTER(BandPlanC), i32, fp, fp, vp, vp, vp]),
        "thz_trace_fused_host": (i32, [vp, fp, fp, fp, i64]),
        "thz_trace_forward_host": (i32, [vp, fp, fp, fp, fp, fp, i64]),
        "thz_trace_inverse_host": (i32, [vp, fp, i32, i32, fp, fp, i64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._signatures = sig
    _lib = L
    return L


class _LazyLib:
    def __getattr__(self, name):
        return getattr(load_library(), name)


lib = _LazyLib()


def _ptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    return a.ctypes.data


def _f32c(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


class DeviceBuffer:
    """A device allocation owned by a Context (thz_dev_alloc / thz_dev_free)."""

    def __init__(self, ctx: "Context", nbytes: int):
        self.ctx = ctx
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        ctx._check(lib.thz_dev_alloc(ctx.handle, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def free(self):
        if self.ptr is not None and self.ctx.handle is not None:
            lib.thz_dev_free(self.ctx.handle, self.ptr)
        self.ptr = None

    def upload(self, a: np.ndarray):
        a = np.ascontiguousarray(a)
        assert a.nbytes <= self.nbytes
        self.ctx._check(lib.thz_copy_h2d(self.ctx.handle, self.ptr, a.ctypes.data, a.nbytes))
        return self

    def download(self, shape, dtype=np.float32, offset_bytes=0):
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes + offset_bytes <= self.nbytes
        self.ctx._check(lib.thz_copy_d2h(self.ctx.handle, out.ctypes.data, self.ptr + offset_bytes, out.nbytes))
        return out

    def zero(self):
        self.ctx._check(lib.thz_dev_memset(self.ctx.handle, self.ptr, 0, self.nbytes))
        return self

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One GPU's libthzgpu context (thz_ctx_create)."""

    def __init__(self, device: int = 0):
        self.handle = None
        h = C.c_void_p()
        rc = lib.thz_ctx_create(int(device), C.byref(h))
        if rc != THZ_OK:
            raise ThzError(rc, (lib.thz_last_error(None) or b"").decode())
        self.handle = h.value
        self.n = 0

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc not in (THZ_OK, THZ_ABORTED):
            raise ThzError(rc, (lib.thz_last_error(self.handle) or b"").decode())
        return rc

    def close(self):
        if self.handle is not None:
            lib.thz_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def sm_count(self):
        return lib.thz_ctx_sm_count(self.handle)

    @property
    def stream(self):
        return lib.thz_ctx_stream(self.handle)

    @property
    def launches(self):
        return int(lib.thz_launch_count(self.handle))

    def sync(self):
        self._check(lib.thz_sync(self.handle))

    def alloc(self, nbytes) -> DeviceBuffer:
        return DeviceBuffer(self, nbytes)

    def to_device(self, a: np.ndarray) -> DeviceBuffer:
        a = np.ascontiguousarray(a)
        return DeviceBuffer(self, max(a.nbytes, 16)).upload(a)

    def generate_cube(self, buf: DeviceBuffer, width, height, n, row0=0, total_width=None, seed=20261018,
                      t0=1000.0, dt=0.05, noise=0.01):
        self._check(lib.thz_generate_cube(self.handle, buf.ptr, width, height, n, row0,
                                          total_width if total_width is not None else width, seed, t0, dt, noise))

    # ------------------------------------------------------------------ trace plan
    def plan_trace(self, n, m_pre=None, band=None, m_post=None):
        """thz_plan_trace: multiplier vectors are plain float32 arrays computed by the caller
        (host.py computes them through the library's own host routines)."""
        keep = []

        def vec(a, ln):
            if a is None:
                return None
            a = _f32c(a, (ln,))
            keep.append(a)
            return a.ctypes.data

        self._check(lib.thz_plan_trace(self.handle, int(n), vec(m_pre, n), vec(band, n // 2 + 1), vec(m_post, n)))
        self.n = int(n)

    # ------------------------------------------------------------------ device-pointer ops
    def trace_fused_dev(self, d_in, d_out, d_img, P):
        self._check(lib.thz_trace_fused_dev(self.handle, d_in, d_out, d_img, int(P)))

    def trace_forward_dev(self, d_in, d_win, d_fft, d_amp, d_phase, P):
        self._check(lib.thz_trace_forward_dev(self.handle, d_in, d_win, d_fft, d_amp, d_phase, int(P)))

    def trace_inverse_dev(self, d_fft, use_band, use_post, d_out, d_img, P):
        self._check(lib.thz_trace_inverse_dev(self.handle, d_fft, int(use_band), int(use_post), d_out, d_img, int(P)))

    def band_apply_dev(self, d_fft, d_amp, P):
        self._check(lib.thz_band_apply_dev(self.handle, d_fft, d_amp, int(P)))

    def spectral_means(self, d_fft, d_amp, d_phase, P):
        F = self.n // 2 + 1
        a_fft = np.empty(2 * F, np.float32)
        a_amp = np.empty(F, np.float32)
        a_ph = np.empty(F, np.float32)
        self._check(lib.thz_spectral_means(self.handle, d_fft, d_amp, d_phase, int(P), a_fft.ctypes.data,
                                           a_amp.ctypes.data if d_amp else None,
                                           a_ph.ctypes.data if d_phase else None))
        return a_fft.view(np.complex64), a_amp, a_ph

    # ------------------------------------------------------------------ host-pointer ops
    def trace_fused(self, data: np.ndarray, want_img=True):
        """Slots 2..7 of the default chain on a host cube [..., N] -> (filtered, img)."""
        data = _f32c(data)
        n = data.shape[-1]
        assert n == self.n, "plan_trace(n) first"
        P = data.size // n if n else 0
        out = np.empty_like(data)
        img = np.empty(data.shape[:-1], np.float32) if want_img else None
        self._check(lib.thz_trace_fused_host(self.handle, data.ctypes.data, out.ctypes.data, _ptr(img), P))
        return out, img

    def trace_forward(self, data: np.ndarray, want=("windowed", "fft", "amp", "phase")):
        """`math_tools::fft` on a host cube -> dict(windowed, fft, amp, phase)."""
        data = _f32c(data)
        n = data.shape[-1]
        assert n == self.n, "plan_trace(n) first"
        P = data.size // n if n else 0
        F = n // 2 + 1
        lead = data.shape[:-1]
        res = {}
        if "windowed" in want:
            res["windowed"] = np.empty_like(data)
        if "fft" in want:
            res["fft"] = np.empty(lead + (F,), np.complex64)
        if "amp" in want:
            res["amp"] = np.empty(lead + (F,), np.float32)
        if "phase" in want:
            res["phase"] = np.empty(lead + (F,), np.float32)
        self._check(lib.thz_trace_forward_host(self.handle, data.ctypes.data, _ptr(res.get("windowed")),
                                               _ptr(res.get("fft")), _ptr(res.get("amp")), _ptr(res.get("phase")), P))
        return res

    def trace_inverse(self, fft: np.ndarray, use_band=False, use_post=False, want_img=False):
        """`math_tools::ifft` on a host spectrum cube [..., F] complex64 -> (data, img)."""
        fft = np.ascontiguousarray(fft, dtype=np.complex64)
        n = self.n
        F = n // 2 + 1
        assert fft.shape[-1] == F
        P = fft.size // F
        out = np.empty(fft.shape[:-1] + (n,), np.float32)
        img = np.empty(fft.shape[:-1], np.float32) if want_img else None
        self._check(lib.thz_trace_inverse_host(self.handle, fft.ctypes.data, int(use_band), int(use_post),
                                               out.ctypes.data, _ptr(img), P))
        return out, img

    # ------------------------------------------------------------------ deconvolution
    def deconv_energies_dev(self, d_cube, P, n, bands, d_energy):
        self._check(lib.thz_deconv_energies_dev(self.handle, d_cube, int(P), int(n), bands, len(bands), d_energy))

    def chain_energies_dev(self, d_in, d_out, d_img, P, n, bands, d_energy):
        """thz_chain_energies_dev: trace pass fused with the band-energy pass."""
        self._check(lib.thz_chain_energies_dev(self.handle, d_in, d_out, d_img, int(P), int(n), bands, len(bands),
                                               d_energy))

    def chain_begin_dev(self, d_in, d_work, d_img, P, n, bands, d_energy):
        """thz_chain_begin_dev: trace pass + band energies, hand-off for thz_chain_end_dev left in d_work."""
        self._check(lib.thz_chain_begin_dev(self.handle, d_in, d_work, d_img, int(P), int(n), bands, len(bands),
                                            d_energy))

    def chain_end_dev(self, d_work, d_gain, P, n, bands, d_out, d_img):
        self._check(lib.thz_chain_end_dev(self.handle, d_work, d_gain, int(P), int(n), bands, len(bands), d_out, d_img))

    def chain_dev(self, d_in, rows, cols, n, bands, d_out, d_img):
        """thz_chain_dev: default chain + deconvolution on a device-resident cube."""
        self._check(lib.thz_chain_dev(self.handle, d_in, int(rows), int(cols), int(n), bands, len(bands), d_out, d_img,
                                      None, None, None))

    def chain_kernel_ms(self):
        ms = np.zeros(5, np.float32)
        self._check(lib.thz_chain_kernel_ms(self.handle, ms.ctypes.data))
        return {"energy_spectra_ms": float(ms[0]), "energy_edges_ms": float(ms[1]), "apply_edges_ms": float(ms[2]),
                "apply_main_ms": float(ms[3]), "trace_ms": float(ms[4])}

    def deconv_apply_dev(self, d_cube, d_gain, P, n, bands, d_out, d_img):
        self._check(lib.thz_deconv_apply_dev(self.handle, d_cube, d_gain, int(P), int(n), bands, len(bands), d_out,
                                             d_img))

    def conv2d(self, image, psf_x=None, psf_y=None, dense=None, direct=True):
        """One 'same' zero-boundary 2-D filtering (test hook of the RL tile kernel)."""
        image = _f32c(image)
        rows, cols = image.shape
        d_in = self.to_device(image)
        d_out = self.alloc(image.nbytes)
        if dense is None:
            px, py = _f32c(psf_x), _f32c(psf_y)
            self._check(lib.thz_conv2d_separable_dev(self.handle, d_in.ptr, rows, cols, px.ctypes.data, px.size,
                                                     py.ctypes.data, py.size, int(direct), d_out.ptr))
        else:
            k = _f32c(dense)
            self._check(lib.thz_conv2d_dense_dev(self.handle, d_in.ptr, rows, cols, k.ctypes.data, k.shape[0],
                                                 k.shape[1], int(direct), d_out.ptr))
        return d_out.download((rows, cols))

    def richardson_lucy(self, image, n_iter, psf_x=None, psf_y=None, dense=None, direct=True, want_gain=False):
        """`richardson_lucy` + clamp (+ gain) on a host image."""
        image = _f32c(image)
        rows, cols = image.shape
        d_in = self.to_device(image)
        d_u = self.alloc(image.nbytes)
        d_g = self.alloc(image.nbytes) if want_gain else None
        if dense is None:
            px, py = _f32c(psf_x), _f32c(psf_y)
            rc = lib.thz_rl_separable_dev(self.handle, d_in.ptr, rows, cols, px.ctypes.data, px.size, py.ctypes.data,
                                          py.size, int(direct), int(n_iter), d_u.ptr, d_g.ptr if d_g else None,
                                          None, None, None, 0.0, 0.0)
        else:
            k = _f32c(dense)
            rc = lib.thz_rl_dense_dev(self.handle, d_in.ptr, rows, cols, k.ctypes.data, k.shape[0], k.shape[1],
                                      int(direct), int(n_iter), d_u.ptr, d_g.ptr if d_g else None, None)
        self._check(rc)
        u = d_u.download((rows, cols))
        return (u, d_g.download((rows, cols))) if want_gain else u

    def deconvolution(self, cube, bands, abort_flag=None, progress=None):
        """`Deconvolution::filter` on a host cube [rows][cols][n] -> (out, img, status)."""
        cube = _f32c(cube)
        rows, cols, n = cube.shape
        out = np.empty_like(cube)
        img = np.empty((rows, cols), np.float32)
        cb = PROGRESS_FN(progress) if progress is not None else None
        rc = self._check(lib.thz_deconvolution_host(
            self.handle, cube.ctypes.data, rows, cols, n, bands, len(bands), out.ctypes.data, img.ctypes.data,
            C.addressof(abort_flag) if abort_flag is not None else None,
            C.cast(cb, C.c_void_p) if cb is not None else None, None))
        return out, img, rc

    def chain(self, cube, bands=None):
        """thz_chain_host: default chain (+ deconvolution when bands is given) on a host cube."""
        cube = _f32c(cube)
        rows, cols, n = cube.shape
        out = np.empty_like(cube)
        img = np.empty((rows, cols), np.float32)
        rc = self._check(lib.thz_chain_host(self.handle, cube.ctypes.data, rows, cols, n, bands,
                                            len(bands) if bands is not None else 0, out.ctypes.data,
                                            img.ctypes.data, None, None, None))
        return out, img, rc

    def scale_blocks(self, a, scale):
        """`scale_3d` on a host array [width][height][z] (float32 or complex64)."""
        a = np.ascontiguousarray(a)
        cplx = np.iscomplexobj(a)
        f = a.astype(np.complex64).view(np.float32) if cplx else a.astype(np.float32)
        w, h, z = f.shape
        out = np.empty((w // scale, h // scale, z), np.float32)
        self._check(lib.thz_scale_blocks_host(self.handle, f.ctypes.data, w, h, z, int(scale), out.ctypes.data))
        return out.view(np.complex64) if cplx else out

    def bias_subtract(self, cube):
        cube = _f32c(cube)
        n = cube.shape[-1]
        P = cube.size // n
        d = self.to_device(cube)
        d_img = self.alloc(max(P * 4, 16))
        self._check(lib.thz_bias_subtract_dev(self.handle, d.ptr, n, d.ptr, d_img.ptr, P))
        return d.download(cube.shape), d_img.download(cube.shape[:-1])

    def roi_average(self, data, polygon, scaling=1):
        """`average_polygon_roi` of a host array [dim0][dim1][z] (uploaded for the call)."""
        data = _f32c(data)
        d0, d1, z = data.shape
        d = self.to_device(data)
        px = np.ascontiguousarray([p[0] for p in polygon], np.int64)
        py = np.ascontiguousarray([p[1] for p in polygon], np.int64)
        out = np.empty(z, np.float32)
        self._check(lib.thz_roi_average_dev(self.handle, d.ptr, d0, d1, z, px.ctypes.data, py.ctypes.data, px.size,
                                            int(scaling), out.ctypes.data))
        return out

    def reference_pulse(self, scan_time, ref_time, ref_signal, window_type=0, fft_window=(1.0, 7.0)):
        st = np.ascontiguousarray(scan_time, np.float32)
        rt = np.ascontiguousarray(ref_time, np.float32)
        rs = np.ascontiguousarray(ref_signal, np.float32)
        n = st.size
        sig = np.empty(n, np.float32)
        amp = np.empty(n // 2 + 1, np.float32)
        ph = np.empty(n // 2 + 1, np.float32)
        self._check(lib.thz_reference_pulse(self.handle, st.ctypes.data, n, rt.ctypes.data, rs.ctypes.data, rt.size,
                                            int(window_type), float(fft_window[0]), float(fft_window[1]),
                                            sig.ctypes.data, amp.ctypes.data, ph.ctypes.data))
        self.n = n
        return sig, amp, ph

    def voxel_opacity(self, cube, opacity_threshold=0.1, contrast=2.0, sigma=3.0, radius=9, max_instances=2_000_000):
        cube = _f32c(cube)
        n = cube.shape[-1]
        P = cube.size // n
        d = self.to_device(cube)
        d_o = self.alloc(cube.nbytes)
        thr = C.c_float(0.0)
        self._check(lib.thz_voxel_opacity_dev(self.handle, d.ptr, n, P, float(opacity_threshold), float(contrast),
                                              float(sigma), int(radius), int(max_instances), d_o.ptr,
                                              C.cast(C.byref(thr), C.c_void_p)))
        return d_o.download(cube.shape), float(thr.value)

    # ------------------------------------------------------------------ config 2 / GUI hand-off
    def plan_reference(self, ref_amp, ref_phase):
        if ref_amp is None:
            self._check(lib.thz_plan_reference(self.handle, None, None, 0))
            return
        a, p = _f32c(ref_amp), _f32c(ref_phase)
        self._check(lib.thz_plan_reference(self.handle, a.ctypes.data, p.ctypes.data, a.size))

    def trace_forward_normalised(self, data):
        """-> (ratio [.., F], dphase [.., F]) : |s| / max(A_r, 1e-12) and unwrap(arg s) - phi_r per pixel."""
        data = _f32c(data)
        n = data.shape[-1]
        P = data.size // n
        F = n // 2 + 1
        d_in = self.to_device(data)
        d_a, d_p = self.alloc(P * F * 4), self.alloc(P * F * 4)
        self._check(lib.thz_trace_forward_normalised_dev(self.handle, d_in.ptr, None, None, d_a.ptr, d_p.ptr, P))
        shp = data.shape[:-1] + (F,)
        return d_a.download(shp), d_p.download(shp)

    def spectral_slice(self, d_array, F, bin_idx, P):
        d_m = self.alloc(P * 4)
        self._check(lib.thz_spectral_slice_dev(self.handle, d_array, int(F), int(bin_idx), d_m.ptr, int(P)))
        return d_m.download((P,))

    def pixel_handoff(self, d_raw, d_filtered, P, pixel):
        n = self.n
        F = n // 2 + 1
        raw, fil = np.empty(n, np.float32), np.empty(n, np.float32)
        fft, amp, ph = np.empty(2 * F, np.float32), np.empty(F, np.float32), np.empty(F, np.float32)
        self._check(lib.thz_pixel_handoff_dev(self.handle, d_raw, d_filtered, int(P), int(pixel), raw.ctypes.data,
                                              fil.ctypes.data, fft.ctypes.data, amp.ctypes.data, ph.ctypes.data))
        return {"raw": raw, "filtered": fil, "fft": fft.view(np.complex64), "amp": amp, "phase": ph}

    def mean_trace(self, d_cube, n, P):
        avg = np.empty(n, np.float32)
        self._check(lib.thz_mean_trace_dev(self.handle, d_cube, int(n), int(P), avg.ctypes.data))
        return avg

    def mean_spectra(self, d_raw, P):
        F = self.n // 2 + 1
        f, a, p = np.empty(2 * F, np.float32), np.empty(F, np.float32), np.empty(F, np.float32)
        self._check(lib.thz_mean_spectra_dev(self.handle, d_raw, int(P), f.ctypes.data, a.ctypes.data, p.ctypes.data))
        return f.view(np.complex64), a, p

    def fp32_rate(self, mode=0):
        """thz_fp32_rate: lane operations per second of FFMA (0), packed FFMA2 (1), FADD (2), FADD2 (3), FMUL (4), FMUL2 (5)."""
        v = C.c_double()
        self._check(lib.thz_fp32_rate(self.handle, int(mode), C.byref(v)))
        return v.value

    def deconv_stage_ms(self):
        ms = np.zeros(4, np.float32)
        self._check(lib.thz_deconv_stage_ms(self.handle, ms.ctypes.data))
        return {"energies_ms": float(ms[0]), "rl_ms": float(ms[1]), "apply_ms": float(ms[2]), "rl_iterations": int(ms[3])}

    def deconv_kernel_ms(self):
        ms = np.zeros(4, np.float32)
        self._check(lib.thz_deconv_kernel_ms(self.handle, ms.ctypes.data))
        return {"energy_spectra_ms": float(ms[0]), "energy_edges_ms": float(ms[1]), "apply_edges_ms": float(ms[2]),
                "apply_main_ms": float(ms[3])}


class Slab:
    """One rank's share of the row-slab Richardson-Lucy (thz_slab_*): the halo rows travel as peer stores over
    NVLink inside the filtering kernels.  `plan` is collective (same arguments on every rank, all ranks idle)."""

    IPC_BYTES = 64

    def __init__(self, ctx: Context, rank: int, world: int):
        self.ctx, self.rank, self.world = ctx, int(rank), int(world)
        h = C.c_void_p()
        ctx._check(lib.thz_slab_create(ctx.handle, self.rank, self.world, C.byref(h)))
        self.handle = h.value

    def close(self):
        if self.handle and self.ctx.handle is not None:
            lib.thz_slab_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def plan(self, row_bounds, cols, bands) -> int:
        rb = (C.c_int * (self.world + 1))(*[int(v) for v in row_bounds])
        ch = C.c_int(0)
        self.ctx._check(lib.thz_slab_plan(self.handle, rb, int(cols), bands, len(bands), C.byref(ch)))
        return ch.value

    def export(self) -> bytes:
        buf = (C.c_ubyte * self.IPC_BYTES)()
        self.ctx._check(lib.thz_slab_export(self.handle, C.addressof(buf)))
        return bytes(buf)

    def connect_ipc(self, handles_by_rank):
        blob = b"".join(handles_by_rank)
        assert len(blob) == self.world * self.IPC_BYTES
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self.ctx._check(lib.thz_slab_connect_ipc(self.handle, C.addressof(buf)))

    def connect_local(self, up: "Slab | None", down: "Slab | None"):
        self.ctx._check(lib.thz_slab_connect_local(self.handle, up.handle if up else None, down.handle if down else None))

    def rl(self, d_energy, bstride, d_gain, d_deconv=None):
        self.ctx._check(lib.thz_slab_rl(self.handle, d_energy, int(bstride), d_gain, d_deconv))

    def status(self):
        self.ctx._check(lib.thz_slab_status(self.handle))

    @staticmethod
    def rl_serial(slabs, d_energy, bstride, d_gain, d_deconv=None):
        """All ranks on ONE device, one stream, dependency order (single-GPU emulation of the exchange)."""
        n = len(slabs)
        hs = (C.c_void_p * n)(*[s.handle for s in slabs])
        es = (C.c_void_p * n)(*d_energy)
        gs = (C.c_void_p * n)(*d_gain)
        us = (C.c_void_p * n)(*(d_deconv if d_deconv is not None else [None] * n))
        bs = (C.c_int64 * n)(*[int(b) for b in bstride])
        slabs[0].ctx._check(lib.thz_slab_rl_serial(hs, n, es, bs, gs, us))


class Group:
    """Several GPUs behind one calling thread (thz_group_*): row slabs, halo-exchanged Richardson-Lucy."""

    def __init__(self, devices):
        self.handle = None
        dv = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        rc = lib.thz_group_create(dv, len(devices), C.byref(h))
        if rc != THZ_OK:
            raise ThzError(rc, (lib.thz_last_error(None) or b"").decode())
        self.handle = h.value
        self.size = len(devices)

    def _check(self, rc):
        if rc not in (THZ_OK, THZ_ABORTED):
            raise ThzError(rc, (lib.thz_group_last_error(self.handle) or b"").decode())
        return rc

    def close(self):
        if self.handle is not None:
            lib.thz_group_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def row_bounds(self, rows):
        b = (C.c_int * (self.size + 1))()
        self._check(lib.thz_group_row_bounds(self.handle, int(rows), b))
        return list(b)

    def rl(self, energy, bands):
        """energy [B][rows][cols] host array -> gains of the same shape."""
        e = _f32c(energy)
        nb, rows, cols = e.shape
        assert nb == len(bands)
        g = np.empty_like(e)
        self._check(lib.thz_group_rl_host(self.handle, e.ctypes.data, rows, cols, bands, nb, g.ctypes.data))
        return g

    def chain(self, cube, m_pre=None, band=None, m_post=None, bands=None):
        cube = _f32c(cube)
        rows, cols, n = cube.shape
        out = np.empty_like(cube)
        img = np.empty((rows, cols), np.float32)
        vecs = [None if v is None else _f32c(v) for v in (m_pre, band, m_post)]
        rc = self._check(lib.thz_group_chain_host(
            self.handle, cube.ctypes.data, rows, cols, n, *[_ptr(v) for v in vecs], bands,
            len(bands) if bands is not None else 0, out.ctypes.data, img.ctypes.data, None, None, None))
        return out, img, rc


class Chain:
    """Python handle over the C++ ChainDriver (the mirror of data_thread's chain loop)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        h = C.c_void_p()
        ctx._check(lib.thz_chain_create(ctx.handle, C.byref(h)))
        self.handle = h.value
        self.shape = (0, 0, 0)

    def close(self):
        if self.handle:
            lib.thz_chain_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def stages(self):
        return [lib.thz_chain_stage_name(self.handle, i).decode() for i in range(lib.thz_chain_length(self.handle))]

    def slot_of(self, stage_name):
        return self.stages().index(stage_name) + 1

    def set_param(self, filt, param, value):
        self.ctx._check(lib.thz_chain_set_param(self.handle, filt.encode(), param.encode(), float(value)))

    def get_param(self, filt, param):
        v = C.c_double()
        self.ctx._check(lib.thz_chain_get_param(self.handle, filt.encode(), param.encode(), C.byref(v)))
        return v.value

    def set_active(self, filt, active):
        self.ctx._check(lib.thz_chain_set_active(self.handle, filt.encode(), int(active)))

    def set_config(self, fft_window=(1.0, 7.0), window_type=0, scale_factor=1):
        self.ctx._check(lib.thz_chain_set_config(self.handle, float(fft_window[0]), float(fft_window[1]),
                                                 int(window_type), int(scale_factor)))

    def set_psf(self, psf):
        self.ctx._check(lib.thz_chain_set_psf(self.handle, C.byref(psf.c)))

    def open(self, time, cube, dx=None, dy=None):
        t = np.ascontiguousarray(time, np.float32)
        cube = _f32c(cube)
        w, h, n = cube.shape
        self.shape = (w, h, n)
        self.ctx._check(lib.thz_chain_open(self.handle, t.ctypes.data, n, cube.ctypes.data, w, h,
                                           int(dx is not None and dy is not None), float(dx or 0), float(dy or 0)))

    def run(self, start_idx=1, run_deconvolution=False):
        self.ctx._check(lib.thz_chain_run(self.handle, int(start_idx), int(run_deconvolution)))

    def run_fused(self, run_deconvolution=False):
        self.ctx._check(lib.thz_chain_run_fused(self.handle, int(run_deconvolution)))
        d, i = C.c_void_p(), C.c_void_p()
        lib.thz_chain_fused_result(self.handle, C.byref(d), C.byref(i))
        w, h, n = self.shape
        return _as_np(d.value, (w, h, n)), _as_np(i.value, (w, h))

    def abort(self, value=True):
        lib.thz_chain_abort(self.handle, int(value))

    def filter_ms(self, filt):
        return float(lib.thz_chain_filter_ms(self.handle, filt.encode()))

    def slot(self, idx):
        p = [C.c_void_p() for _ in range(8)]
        n, f = C.c_int(), C.c_int()
        self.ctx._check(lib.thz_chain_slot(self.handle, int(idx), *[C.byref(x) for x in p], C.byref(n), C.byref(f)))
        w, h, _ = self.shape
        N, F = n.value, f.value
        names = ["data", "fft", "amplitudes", "phases", "img", "avg_fft", "avg_signal_fft", "avg_phase_fft"]
        shapes = [(w, h, N), (w, h, F), (w, h, F), (w, h, F), (w, h), (F,), (F,), (F,)]
        out = {}
        for nm, ptr, shp in zip(names, p, shapes):
            if nm in ("fft", "avg_fft"):
                out[nm] = _as_np(ptr.value, shp + (2,)).view(np.complex64).reshape(shp)
            else:
                out[nm] = _as_np(ptr.value, shp)
        return out
