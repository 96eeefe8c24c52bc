"""Host-side mirror of the reference's filter parameters -> multiplier vectors.

Thin wrappers over the library's own C++ host routines (csrc/thz_windows.cpp); nothing is
computed in Python.  Names follow the reference: `FftWindowType` (src/config.rs),
`TimeDomainBandPassBeforeFFT` / `AfterFFT`, `FrequencyDomainBandPass`, `TiltCompensation`.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from .binding import (BandPlanC, DeconvParamsC, HybridFitC, PsfC, SplineC, ThzError, lib, SKIP_REASONS,
                      THZ_FIR_TAPS)

FFT_WINDOW_TYPES = {"AdaptedBlackman": 0, "Blackman": 1, "Hanning": 2, "Hamming": 3, "FlatTop": 4}


def _chk(rc, what):
    if rc != 0:
        raise ThzError(rc, what)


def _t32(time):
    return np.ascontiguousarray(time, dtype=np.float32)


def frequency_axis(time):
    t = _t32(time)
    f = np.empty(t.size // 2 + 1, np.float32)
    _chk(lib.thz_frequency_axis(t.ctypes.data, t.size, f.ctypes.data), "thz_frequency_axis")
    return f


def adapted_blackman(axis, lo, hi):
    a = _t32(axis)
    m = np.empty(a.size, np.float32)
    _chk(lib.thz_adapted_blackman(a.ctypes.data, a.size, float(lo), float(hi), m.ctypes.data), "thz_adapted_blackman")
    return m


def fft_window(time, window_type="AdaptedBlackman", fft_window=(1.0, 7.0)):
    t = _t32(time)
    m = np.empty(t.size, np.float32)
    _chk(lib.thz_window_multiplier(FFT_WINDOW_TYPES[window_type], t.ctypes.data, t.size, float(fft_window[0]),
                                   float(fft_window[1]), m.ctypes.data), "thz_window_multiplier")
    return m


def time_gate(time, low, high, window_width):
    """-> (mult, lower, upper, clamped_low, clamped_high)"""
    t = _t32(time)
    m = np.empty(t.size, np.float32)
    lo, hi = C.c_double(float(low)), C.c_double(float(high))
    lower, upper = C.c_int(), C.c_int()
    _chk(lib.thz_time_gate_multiplier(t.ctypes.data, t.size, C.byref(lo), C.byref(hi), float(window_width),
                                      m.ctypes.data, C.byref(lower), C.byref(upper)), "thz_time_gate_multiplier")
    return m, lower.value, upper.value, lo.value, hi.value


def band_pass(freq, low=0.2, high=5.0, window_width=0.1):
    """-> (mult, lower, upper)"""
    f = _t32(freq)
    m = np.empty(f.size, np.float32)
    lower, upper = C.c_int(), C.c_int()
    _chk(lib.thz_band_pass_multiplier(f.ctypes.data, f.size, float(low), float(high), float(window_width),
                                      m.ctypes.data, C.byref(lower), C.byref(upper)), "thz_band_pass_multiplier")
    return m, lower.value, upper.value


@dataclass
class ChainConfig:
    """Parameters of the default chain with the reference's defaults (src/config.rs:203-212,
    src/filters/band_pass_fd.rs:52-54, band_pass_td_before_fft.rs:52-54, band_pass_td_after_fft.rs:54,
    tilt_compensation.rs:188)."""
    fft_window_type: str = "AdaptedBlackman"
    fft_window: tuple = (1.0, 7.0)
    tilt_active: bool = True            # 0 deg tilt: tapers the last 7 ps when dx/dy are present
    gate_before_active: bool = True
    gate_before_window: float = 2.0
    band_active: bool = True
    band_low: float = 0.2
    band_high: float = 5.0
    band_window: float = 0.1
    gate_after_active: bool = True
    gate_after_window: float = 0.1
    gate_before: tuple = field(default=None)   # (low, high) in ps; None = full axis (`reset`)
    gate_after: tuple = field(default=None)


def chain_multipliers(time, cfg: ChainConfig = None, dx_dy_present=True):
    """(m_pre, band, m_post) for thz_plan_trace: sequential f32 products in chain order
    (tilt taper, gate before FFT, FFT window), like the reference's successive in-place
    multiplications."""
    cfg = cfg or ChainConfig()
    t = _t32(time)
    m_pre = np.ones(t.size, np.float32)
    if cfg.tilt_active and dx_dy_present:
        m_pre = m_pre * adapted_blackman(t, 0.0, 7.0)
    if cfg.gate_before_active:
        lo, hi = cfg.gate_before if cfg.gate_before is not None else (float(t[0]), float(t[-1]))
        m_pre = m_pre * time_gate(t, lo, hi, cfg.gate_before_window)[0]
    m_pre = m_pre * fft_window(t, cfg.fft_window_type, cfg.fft_window)
    f = frequency_axis(t)
    band = band_pass(f, cfg.band_low, cfg.band_high, cfg.band_window)[0] if cfg.band_active else None
    if cfg.gate_after_active:
        lo, hi = cfg.gate_after if cfg.gate_after is not None else (float(t[0]), float(t[-1]))
        m_post = time_gate(t, lo, hi, cfg.gate_after_window)[0]
    else:
        m_post = None
    return m_pre.astype(np.float32), band, m_post


# ----------------------------------------------------------------------------------------
# Deconvolution: PSF container (`load_psf`, src/io.rs:190-267) and the band planner
# ----------------------------------------------------------------------------------------
class PSF:
    """The 26 arrays of psf.npz cast to f32, held alive for the C structs that borrow them."""

    def __init__(self, arrays: dict):
        self._keep = []
        self.c = PsfC()

        def arr(name):
            a = np.ascontiguousarray(np.asarray(arrays[name], dtype=np.float64).reshape(-1).astype(np.float32))
            self._keep.append(a)
            return a

        def spline(prefix, dst: SplineC):
            k = arr(f"{prefix}_knots_thz")
            dst.n = k.size
            dst.knots = k.ctypes.data
            dst.values = arr(f"{prefix}_values_mm").ctypes.data
            for nm in "abcd":
                setattr(dst, f"coeff_{nm}", arr(f"{prefix}_coeff_{nm}").ctypes.data)

        def hybrid(prefix, dst: HybridFitC):
            dst.base_a = float(arr(f"{prefix}_base_a")[0])
            dst.base_b = float(arr(f"{prefix}_base_b")[0])
            spline(f"{prefix}_corr", dst.correction)

        hybrid("wx", self.c.wx_fit)
        hybrid("wy", self.c.wy_fit)
        spline("x0", self.c.x0_spline)
        spline("y0", self.c.y0_spline)

    @classmethod
    def load(cls, path):
        z = np.load(path)
        return cls({k: z[k] for k in z.files})

    def wx(self, f):
        return float(lib.thz_hybrid_eval(C.byref(self.c.wx_fit), float(f)))

    def wy(self, f):
        return float(lib.thz_hybrid_eval(C.byref(self.c.wy_fit), float(f)))

    def x0(self, f):
        return float(lib.thz_spline_eval_const_extrap(C.byref(self.c.x0_spline), float(f)))

    def y0(self, f):
        return float(lib.thz_spline_eval_const_extrap(C.byref(self.c.y0_spline), float(f)))


def fir_bank(n_filters, start_freq, end_freq, win_width, time):
    t = _t32(time)
    filt = np.empty((n_filters, THZ_FIR_TAPS), np.float32)
    cen = np.empty(n_filters, np.float32)
    _chk(lib.thz_fir_bank(int(n_filters), float(start_freq), float(end_freq), float(win_width), float(t[0]),
                          float(t[1]), filt.ctypes.data, cen.ctypes.data), "thz_fir_bank")
    return filt, cen


@dataclass
class Deconvolution:
    """`Deconvolution` filter parameters (src/filters/deconvolution.rs:725-733)."""
    n_iterations: int = 500
    n_filters: int = 25
    start_freq: float = 0.1
    end_freq: float = 10.0
    win_width: float = 0.5

    def plan(self, time, shape, dx, dy, psf: PSF):
        """-> (bands ctypes array, None) or (None, reason) when the reference returns its input."""
        t = _t32(time)
        prm = DeconvParamsC(int(self.n_iterations), int(self.n_filters), float(np.float32(self.start_freq)),
                            float(np.float32(self.end_freq)), float(np.float32(self.win_width)))
        bands = (BandPlanC * int(self.n_filters))()
        has = dx is not None and dy is not None
        rc = lib.thz_deconv_plan_bands(C.byref(psf.c) if psf is not None else None, C.byref(prm), t.ctypes.data,
                                       t.size, int(shape[0]), int(shape[1]), int(has), float(dx or 0.0),
                                       float(dy or 0.0), bands)
        if rc in SKIP_REASONS:
            return None, SKIP_REASONS[rc]
        _chk(rc, "thz_deconv_plan_bands")
        return bands, None


def optical_properties(sample_amp, sample_phase, ref_amp, ref_phase, freqs, thickness):
    """`calculate_optical_properties` -> (n, alpha, kappa), computed by the library's host routine."""
    arrs = [np.ascontiguousarray(a, np.float32) for a in (sample_amp, sample_phase, ref_amp, ref_phase, freqs)]
    f = arrs[4].size
    n, al, ka = (np.empty(f, np.float32) for _ in range(3))
    _chk(lib.thz_optical_properties(*[a.ctypes.data for a in arrs], f, float(thickness), n.ctypes.data,
                                    al.ctypes.data, ka.ctypes.data), "thz_optical_properties")
    return n, al, ka
