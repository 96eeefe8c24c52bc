"""Shared helpers for the parity tests (test infrastructure only)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import thz_oracle as orc  # noqa: E402

F32 = np.float32
# north-star tolerances (BASELINE.json): max|a-b| / max|b|
TOL_TRACE = 1e-4     # f32 spectra and filtered traces
TOL_MAP = 1e-3       # deconvolved maps after N iterations


def pkg():
    return importlib.import_module("thz-image-explorer_b200")


def rel_err(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.size == 0:
        return 0.0
    den = float(np.max(np.abs(b)))
    if den == 0.0:
        return float(np.max(np.abs(a)))
    wide = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    return float(np.max(np.abs(a.astype(wide) - b.astype(wide))) / den)


def time_axis(n, t0=1000.0, dt=0.05):
    return (F32(t0) + F32(dt) * np.arange(n, dtype=F32)).astype(F32)


def synthetic_cube(w, h, n, seed=0, noise=0.01, dt=0.05):
    """THz-like pulses: A exp(-((t-tp)/tau)^2) cos(2 pi fc (t-tp)) + noise (numpy restatement of
    the shape of thz_generate_cube; values are not meant to be bit-identical to the device RNG)."""
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) * dt
    amp = rng.uniform(0.5, 1.5, size=(w, h, 1))
    tp = min(10.0, 0.4 * n * dt) + min(5.0, 0.2 * n * dt) * rng.uniform(size=(w, h, 1))
    tt = t[None, None, :] - tp
    x = amp * np.exp(-(tt / 0.3) ** 2) * np.cos(2 * np.pi * 1.0 * tt) + noise * rng.standard_normal((w, h, n))
    return x.astype(F32)


def check_unwrapped_phase(phase_gpu, phase_ref, fft_ref, tol_rad):
    """Phases agree up to threshold-unwrap decisions that are legitimately ambiguous.

    The reference unwrap (src/math_tools.rs:224-237) adds -+2 pi when a raw phase step
    exceeds pi.  Two correct f32 FFTs differ by ~1e-7 of the spectrum peak, so on a bin
    whose raw step is within `amb` of +-pi (or whose amplitude is ~0, where atan2 is
    ill-conditioned) the decision may flip, shifting all later bins by 2 pi.  Accept exactly that:
    (a) phase_gpu - phase_ref is within tol_rad of an integer multiple of 2 pi except on
        ill-conditioned bins, and
    (b) the integer changes only at ambiguous bins.
    Returns the number of flips accepted."""
    pg = np.asarray(phase_gpu, np.float64)
    pr = np.asarray(phase_ref, np.float64)
    z = np.asarray(fft_ref)
    assert pg.shape == pr.shape
    two_pi = 2 * np.pi
    diff = pg - pr
    k = np.round(diff / two_pi)
    resid = np.abs(diff - k * two_pi)
    amp = np.abs(z)
    peak = amp.max(axis=-1, keepdims=True)
    # phase error of a bin ~ (fft abs error ~ 2e-6 * peak) / amp
    cond = 4e-6 * peak / np.maximum(amp, 1e-30)
    ill = cond > 0.25 * tol_rad
    bad = (resid > tol_rad + cond) & ~ill
    assert not bad.any(), f"{bad.sum()} phase bins off by more than {tol_rad} rad (max {resid[~ill].max()})"
    raw = np.arctan2(z.imag.astype(np.float64), z.real.astype(np.float64))
    step = np.abs(np.diff(raw, axis=-1))
    amb_width = tol_rad + cond[..., 1:] + cond[..., :-1]
    ambiguous = (np.abs(step - np.pi) < amb_width) | ill[..., 1:] | ill[..., :-1]
    # on ill-conditioned bins k itself is meaningless; propagate the last well-conditioned k
    dk = np.diff(k, axis=-1) != 0
    illegal = dk & ~ambiguous
    # a change right after an ill-conditioned stretch is attributed to that stretch
    assert not illegal.any(), f"{illegal.sum()} unwrap decisions differ on unambiguous bins"
    return int(dk.sum())


def default_multipliers(n, dx_dy_present=True):
    t = time_axis(n)
    return (t,) + orc.default_chain_multipliers(t, dx_dy_present=dx_dy_present)


def slot0(cube, t, dx=0.5, dy=0.5):
    """A ScannedImageFilterData as the reference holds it after loading (no bias subtraction:
    the cube is taken as already loaded)."""
    w, h, n = cube.shape
    f = orc.frequency_axis(t)
    return orc.ScannedImageFilterData(
        time=t, data=cube.copy(), frequency=f,
        fft=np.zeros((w, h, f.shape[0]), np.complex64), amplitudes=np.zeros((w, h, f.shape[0]), F32),
        phases=np.zeros((w, h, f.shape[0]), F32), img=orc.intensity_image(cube), dx=dx, dy=dy, width=w, height=h)
