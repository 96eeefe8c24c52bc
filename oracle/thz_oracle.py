"""CPU oracle for the thz-image-explorer filter-chain hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker or the timed CPU baseline.  The product
path (``thz-image-explorer_b200``) never imports this module.

It is a line-by-line numpy restatement of the reference's Rust CPU path
(citations are relative to ``/root/reference``):

* windows / unwrap / fft / ifft / scaling ....... ``src/math_tools.rs:81-571``
* frequency band-pass ........................... ``src/filters/band_pass_fd.rs:122-220``
* time gates (before FFT / after iFFT) .......... ``src/filters/band_pass_td_before_fft.rs:124-182``
* tilt compensation ............................. ``src/filters/tilt_compensation.rs:97-226``
* FIR bank, convolve1d, convolve2d, RL, filter .. ``src/filters/deconvolution.rs:30-1041``
* PSF model ..................................... ``src/filters/psf.rs:26-332``
* psf.npz loader ................................ ``src/io.rs:190-267``
* load-time bias subtraction, frequency axis .... ``src/io.rs:578-620``
* intensity image ............................... ``src/data_thread.rs:1288-1307``

PARITY PINNING.  The Rust reference cannot be compiled in this environment (no
rustc/cargo) and its FFT arithmetic lives in un-vendored crates (realfft 3.5.0 on
rustfft 6.4.1, ``Cargo.lock:6507,6747``); their published conventions are restated
here (forward and inverse both unnormalised, N/2+1 bins, the caller divides by N).
The oracle is pinned against every known-answer test the reference holds for this
path (``tests/test_oracle_reference_tests.py`` re-creates them: window end values /
symmetry, FFT round trip 1e-4, exact zeros outside the FD and TD pass-bands, tilt
extension and impulse position, deconvolution shape preservation with the real
``sample_data/psf.npz``).  Those tests pin *properties*, not absolute spectra, phases,
FIR outputs, PSFs or RL iterates: for those values the oracle is "parity unpinned"
beyond the source text, exactly as SURVEY.md §8(c) states.

``dtype=np.float32`` reproduces the reference's arithmetic type; ``np.float64`` is the
ground-truth mode used to bound both implementations' rounding error.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import scipy.fft as sfft

F32 = np.float32
PI32 = np.float32(math.pi)  # std::f32::consts::PI


# --------------------------------------------------------------------------------------
# helpers that emulate Rust f32 libm calls (cos/exp are computed in double and rounded,
# which is what glibc's cosf/expf effectively deliver: < 1 ulp, almost always correctly
# rounded)
# --------------------------------------------------------------------------------------
def _cos32(x):
    return np.cos(np.asarray(x, dtype=np.float64)).astype(F32)


def _exp32(x):
    return np.exp(np.asarray(x, dtype=np.float64)).astype(F32)


# --------------------------------------------------------------------------------------
# math_tools.rs windows
# --------------------------------------------------------------------------------------
def blackman_window(n, m):
    """`blackman_window` (src/math_tools.rs:81-90). f32 arithmetic, NaN -> 1, clamp [0,1]."""
    n = np.asarray(n, dtype=F32)
    m = np.asarray(m, dtype=F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        a = (F32(2.0) * PI32 * n) / m
        b = (F32(4.0) * PI32 * n) / m
        res = F32(0.42) - F32(0.5) * _cos32(a) + F32(0.08) * _cos32(b)
    res = np.asarray(res, dtype=F32)
    out = np.clip(res, F32(0.0), F32(1.0))
    out = np.where(np.isnan(res), F32(1.0), out)
    return out.astype(F32)


def adapted_blackman_multiplier(time, lower_bound, upper_bound):
    """The per-sample factor `apply_adapted_blackman_window` multiplies a trace by
    (src/math_tools.rs:102-122).  Returns a vector of len(time): bw on the tapered
    samples, exactly 1 elsewhere (the reference does not touch those samples)."""
    t = np.asarray(time, dtype=F32)
    lo = F32(lower_bound)
    hi = F32(upper_bound)
    mult = np.ones(t.shape[0], dtype=F32)
    if t.shape[0] == 0:
        return mult
    t0 = t[0]
    tl = t[-1]
    first = t <= (lo + t0)
    second = (~first) & (t >= (tl - hi))
    if first.any():
        mult[first] = blackman_window(t[first] - t0, F32(2.0) * lo)
    if second.any():
        mult[second] = blackman_window(t[second] - (tl - hi * F32(2.0)), F32(2.0) * hi)
    return mult


def apply_adapted_blackman_window(signal, time, lower_bound, upper_bound):
    """In place on the last axis of `signal` (src/math_tools.rs:102-122)."""
    signal *= adapted_blackman_multiplier(time, lower_bound, upper_bound).astype(signal.dtype)
    return signal


def normalize_time(time):
    """src/math_tools.rs:131-135."""
    t = np.asarray(time, dtype=F32)
    mn = t.min()
    mx = t.max()
    return ((t - mn) / (mx - mn)).astype(F32)


def hamming_multiplier(time):
    """src/math_tools.rs:145-150."""
    t = normalize_time(time)
    return (F32(0.54) - F32(0.46) * _cos32(F32(2.0) * PI32 * t)).astype(F32)


def hanning_multiplier(time):
    """src/math_tools.rs:160-165."""
    t = normalize_time(time)
    return (F32(0.5) * (F32(1.0) - _cos32(F32(2.0) * PI32 * t))).astype(F32)


def blackman_multiplier(time):
    """src/math_tools.rs:174-180."""
    t = normalize_time(time)
    return (F32(0.42) - F32(0.5) * _cos32(F32(2.0) * PI32 * t)
            + F32(0.08) * _cos32(F32(4.0) * PI32 * t)).astype(F32)


def flat_top_multiplier(time):
    """src/math_tools.rs:190-198."""
    t = normalize_time(time)
    return (F32(1.0) - F32(1.93) * _cos32(F32(2.0) * PI32 * t)
            + F32(1.29) * _cos32(F32(4.0) * PI32 * t)
            - F32(0.388) * _cos32(F32(6.0) * PI32 * t)
            + F32(0.028) * _cos32(F32(8.0) * PI32 * t)).astype(F32)


WINDOW_TYPES = ("AdaptedBlackman", "Blackman", "Hanning", "Hamming", "FlatTop")


def fft_window_multiplier(time, window_type="AdaptedBlackman", fft_window=(1.0, 7.0)):
    """The window `fft()` applies, as a multiplier vector (src/math_tools.rs:356-371)."""
    if window_type == "AdaptedBlackman":
        return adapted_blackman_multiplier(time, fft_window[0], fft_window[1])
    if window_type == "Blackman":
        return blackman_multiplier(time)
    if window_type == "Hanning":
        return hanning_multiplier(time)
    if window_type == "Hamming":
        return hamming_multiplier(time)
    if window_type == "FlatTop":
        return flat_top_multiplier(time)
    raise ValueError(window_type)


# --------------------------------------------------------------------------------------
# numpy_unwrap (src/math_tools.rs:211-240) -- threshold form, sequential running sum
# --------------------------------------------------------------------------------------
def numpy_unwrap(x, period=None):
    """Sequential unwrap along the last axis. Vectorised over leading axes only."""
    x = np.asarray(x)
    dt = x.dtype.type
    if period is None:
        diff = x[..., 1:] - x[..., :-1]
        period = dt(2.0) * dt(math.pi) / (diff.sum(axis=-1, dtype=x.dtype) / dt(diff.shape[-1]))
    period = dt(period)
    half = period / dt(2.0)
    out = x.copy()
    prev_val = x[..., 0].copy()
    prev_unwrapped = x[..., 0].copy()
    for i in range(1, x.shape[-1]):
        val = x[..., i]
        diff = val - prev_val
        diff = np.where(diff > half, diff - period, np.where(diff < -half, diff + period, diff))
        unwrapped_val = prev_unwrapped + diff
        prev_val = val
        prev_unwrapped = unwrapped_val
        out[..., i] = unwrapped_val
    return out


# --------------------------------------------------------------------------------------
# data container (src/data_container.rs:109-162) -- only the fields on the hot path
# --------------------------------------------------------------------------------------
@dataclass
class ConfigContainer:
    """src/config.rs:171-212 (defaults :203-212)."""
    fft_window: Sequence[float] = (1.0, 7.0)
    fft_window_type: str = "AdaptedBlackman"
    scale_factor: int = 1
    avg_in_fourier_space: bool = False


@dataclass
class ScannedImageFilterData:
    """Subset of `ScannedImageFilterData` (src/data_container.rs:109-162).
    data: (x, y, t) C-order; fft/amplitudes/phases: (x, y, f)."""
    time: np.ndarray = field(default_factory=lambda: np.zeros(0, F32))
    data: np.ndarray = field(default_factory=lambda: np.zeros((0, 0, 0), F32))
    frequency: np.ndarray = field(default_factory=lambda: np.zeros(0, F32))
    fft: np.ndarray = field(default_factory=lambda: np.zeros((0, 0, 0), np.complex64))
    amplitudes: np.ndarray = field(default_factory=lambda: np.zeros((0, 0, 0), F32))
    phases: np.ndarray = field(default_factory=lambda: np.zeros((0, 0, 0), F32))
    img: np.ndarray = field(default_factory=lambda: np.zeros((0, 0), F32))
    avg_fft: np.ndarray = field(default_factory=lambda: np.zeros(0, np.complex64))
    avg_signal_fft: np.ndarray = field(default_factory=lambda: np.zeros(0, F32))
    avg_phase_fft: np.ndarray = field(default_factory=lambda: np.zeros(0, F32))
    dx: Optional[float] = None
    dy: Optional[float] = None
    width: int = 0
    height: int = 0
    scaling: int = 1
    has_plan: bool = True  # r2c / c2r present

    def clone(self):
        return ScannedImageFilterData(
            time=self.time.copy(), data=self.data.copy(), frequency=self.frequency.copy(),
            fft=self.fft.copy(), amplitudes=self.amplitudes.copy(), phases=self.phases.copy(),
            img=self.img.copy(), avg_fft=self.avg_fft.copy(),
            avg_signal_fft=self.avg_signal_fft.copy(), avg_phase_fft=self.avg_phase_fft.copy(),
            dx=self.dx, dy=self.dy, width=self.width, height=self.height, scaling=self.scaling,
            has_plan=self.has_plan)


def frequency_axis(time):
    """`f[i] = i / (t[N-1] - t[0])`, F = N/2+1 (src/io.rs:614-620)."""
    t = np.asarray(time, dtype=F32)
    n = t.shape[0]
    rng = t[n - 1] - t[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        return (np.arange(n // 2 + 1, dtype=F32) / rng).astype(F32)


def intensity_image(data):
    """`img[x,y] = sum_t data^2` (src/data_thread.rs:1288-1307; io.rs:587-594)."""
    d = np.asarray(data)
    return np.sum(d * d, axis=2, dtype=d.dtype)


def load_scan(time, raw, dx=None, dy=None, dtype=F32):
    """What `open_scan_from_thz` does to a cube once it is in memory
    (src/io.rs:576-628): per-trace bias subtraction (first sample), intensity image,
    frequency axis, zeroed spectral cubes."""
    raw = np.asarray(raw, dtype=dtype)
    w, h, n = raw.shape
    data = raw - raw[:, :, :1]
    t = np.asarray(time, dtype=F32)
    f = frequency_axis(t)
    return ScannedImageFilterData(
        time=t, data=data, frequency=f,
        fft=np.zeros((w, h, f.shape[0]), np.complex64 if dtype == F32 else np.complex128),
        amplitudes=np.zeros((w, h, f.shape[0]), dtype),
        phases=np.zeros((w, h, f.shape[0]), dtype),
        img=intensity_image(data), dx=dx, dy=dy, width=w, height=h)


# --------------------------------------------------------------------------------------
# scaling (src/math_tools.rs:242-310)
# --------------------------------------------------------------------------------------
def _scale_3d(a, new_w, new_h, s):
    """Block sum in the reference's accumulation order (i outer, j inner), then / s^2."""
    acc = np.zeros((new_w, new_h, a.shape[2]), dtype=a.dtype)
    for i in range(s):
        for j in range(s):
            acc = acc + a[i:new_w * s:s, j:new_h * s:s, :]
    real_t = a.real.dtype.type
    if np.iscomplexobj(acc):
        # num_complex `Complex<f32> / f32` divides the two components; numpy's complex / real goes
        # through the complex division formula and can differ by 1 ulp
        out = np.empty_like(acc)
        out.real = acc.real / real_t(s * s)
        out.imag = acc.imag / real_t(s * s)
        return out.astype(a.dtype)
    return (acc / real_t(s * s)).astype(a.dtype)


def scaling(inp: ScannedImageFilterData, config: ConfigContainer) -> ScannedImageFilterData:
    s = int(config.scale_factor)
    if s <= 1:
        return inp.clone()
    new_w = inp.width // s
    new_h = inp.height // s
    if new_w == 0 or new_h == 0:
        return inp.clone()
    out = inp.clone()
    out.width, out.height, out.scaling = new_w, new_h, s
    if out.dx is not None:
        out.dx = float(F32(out.dx) * F32(s))
    if out.dy is not None:
        out.dy = float(F32(out.dy) * F32(s))
    out.data = _scale_3d(inp.data, new_w, new_h, s)
    out.amplitudes = _scale_3d(inp.amplitudes, new_w, new_h, s)
    out.phases = _scale_3d(inp.phases, new_w, new_h, s)
    out.fft = _scale_3d(inp.fft, new_w, new_h, s)
    return out


# --------------------------------------------------------------------------------------
# fft / ifft (src/math_tools.rs:330-571)
# --------------------------------------------------------------------------------------
def _complex_of(dtype):
    return np.complex64 if np.dtype(dtype) == np.dtype(F32) else np.complex128


def rfft_unnormalised(x, workers=None):
    """realfft `RealToComplex::process` convention: unnormalised, N/2+1 bins
    (call site src/math_tools.rs:375).  pocketfft computes in the input precision."""
    return sfft.rfft(x, axis=-1, workers=workers)


def irfft_unnormalised(spec, n, workers=None):
    """realfft `ComplexToReal::process`: unnormalised inverse (src/math_tools.rs:561);
    imaginary parts of DC / Nyquist are ignored."""
    return sfft.irfft(spec, n=n, axis=-1, norm="forward", workers=workers)


def fft(inp: ScannedImageFilterData, config: ConfigContainer, workers=None) -> ScannedImageFilterData:
    """`fft` (src/math_tools.rs:330-398): window (stored back into data) -> r2c ->
    spectrum, |s|, unwrap(arg s)."""
    out = inp.clone()
    if not out.has_plan:
        return out
    dt = out.data.dtype
    mult = fft_window_multiplier(out.time, config.fft_window_type, config.fft_window).astype(dt)
    out.data = (out.data * mult).astype(dt)
    spec = rfft_unnormalised(out.data, workers=workers)
    out.fft = spec.astype(_complex_of(dt))
    out.amplitudes = np.abs(out.fft).astype(dt)
    phase = np.arctan2(out.fft.imag, out.fft.real).astype(dt)
    out.phases = numpy_unwrap(phase, dt.type(2.0) * dt.type(PI32)).astype(dt)
    return out


def _mean_axis0_twice(a):
    """`mean_axis(Axis(0))` twice (src/math_tools.rs:421-440): sequential sums in the
    array's own precision over x, then over y."""
    acc = np.zeros(a.shape[1:], dtype=a.dtype)
    for i in range(a.shape[0]):
        acc = acc + a[i]
    real_t = a.real.dtype.type
    acc = acc / real_t(a.shape[0])
    acc2 = np.zeros(a.shape[2:], dtype=a.dtype)
    for j in range(acc.shape[0]):
        acc2 = acc2 + acc[j]
    return (acc2 / real_t(acc.shape[0])).astype(a.dtype)


def ifft(inp: ScannedImageFilterData, config: ConfigContainer, workers=None) -> ScannedImageFilterData:
    """`ifft` (src/math_tools.rs:418-571), ROI processing omitted (no ROIs on the hot
    path): pixel means of fft/amplitudes/phases, then c2r / N per trace."""
    out = inp.clone()
    out.avg_fft = _mean_axis0_twice(out.fft)
    out.avg_signal_fft = _mean_axis0_twice(out.amplitudes)
    out.avg_phase_fft = _mean_axis0_twice(out.phases)
    if out.has_plan:
        n = out.time.shape[0]
        dt = out.data.dtype
        y = irfft_unnormalised(out.fft, n, workers=workers).astype(dt)
        out.data = (y / dt.type(n)).astype(dt)
    return out


# --------------------------------------------------------------------------------------
# FrequencyDomainBandPass (src/filters/band_pass_fd.rs:47-58, 122-220)
# --------------------------------------------------------------------------------------
def fd_band_indices(frequency, low, high):
    """[lower, upper) exactly as src/filters/band_pass_fd.rs:135-152."""
    f = np.asarray(frequency, dtype=F32)
    safe_low = F32(max(float(low), 0.0))
    last = float(f[-1]) if f.shape[0] else 10.0
    safe_high = F32(min(float(high), last))
    ge = np.nonzero(f >= safe_low)[0]
    lower = int(ge[0]) if ge.size else 0
    le = np.nonzero(f <= safe_high)[0]
    upper = int(le[-1]) + 1 if le.size else int(f.shape[0])
    return lower, upper


def fd_band_multiplier(frequency, low=0.2, high=5.0, window_width=0.1):
    """The real multiplier over all F bins that the FD band-pass amounts to: adapted
    Blackman taper on [lower, upper), zero elsewhere (band_pass_fd.rs:155-212)."""
    f = np.asarray(frequency, dtype=F32)
    lower, upper = fd_band_indices(f, low, high)
    mult = np.zeros(f.shape[0], dtype=F32)
    if upper > lower:
        mult[lower:upper] = adapted_blackman_multiplier(f[lower:upper], F32(window_width), F32(window_width))
    return mult


@dataclass
class FrequencyDomainBandPass:
    low: float = 0.2
    high: float = 5.0
    window_width: float = 0.1

    def filter(self, inp: ScannedImageFilterData) -> ScannedImageFilterData:
        out = inp.clone()
        mult = fd_band_multiplier(inp.frequency, self.low, self.high, self.window_width)
        rdt = inp.amplitudes.dtype
        out.fft = (inp.fft * mult.astype(rdt)).astype(inp.fft.dtype)
        out.amplitudes = (inp.amplitudes * mult.astype(rdt)).astype(rdt)
        return out


# --------------------------------------------------------------------------------------
# TimeDomainBandPassBeforeFFT / AfterFFT (src/filters/band_pass_td_before_fft.rs:47-72,
# 124-182; band_pass_td_after_fft.rs differs only in window_width default 0.1, :54)
# --------------------------------------------------------------------------------------
def td_gate_indices(time, low, high):
    """[lower, upper) as band_pass_td_before_fft.rs:134-152; returns clamped low/high too."""
    t = np.asarray(time, dtype=F32)
    min_time = float(t[0]) if t.shape[0] else 0.0
    max_time = float(t[-1]) if t.shape[0] else 0.0
    low = max(float(low), min_time)
    high = min(float(high), max_time)
    ge = np.nonzero(t >= F32(low))[0]
    lower = int(ge[0]) if ge.size else 0
    ge = np.nonzero(t >= F32(high))[0]
    upper = int(ge[0]) if ge.size else max(t.shape[0] - 1, 0)
    upper = min(max(upper, lower + 1), t.shape[0])
    return lower, upper, low, high


def td_gate_multiplier(time, low, high, window_width):
    """Per-sample multiplier of the time gate: zero outside [lower, upper), adapted
    Blackman (ww, ww) on the sub-axis time[lower:upper] inside."""
    t = np.asarray(time, dtype=F32)
    lower, upper, _, _ = td_gate_indices(t, low, high)
    mult = np.zeros(t.shape[0], dtype=F32)
    mult[lower:upper] = adapted_blackman_multiplier(t[lower:upper], F32(window_width), F32(window_width))
    return mult


@dataclass
class TimeDomainBandPass:
    low: float = 0.0
    high: float = 0.0
    window_width: float = 2.0  # 2.0 before FFT, 0.1 after iFFT

    def reset(self, time, shape=None):
        """band_pass_td_before_fft.rs:66-72."""
        t = np.asarray(time, dtype=F32)
        self.low = float(t[0]) if t.shape[0] else 0.0
        self.high = float(t[-1]) if t.shape[0] else 0.0

    def filter(self, inp: ScannedImageFilterData) -> ScannedImageFilterData:
        out = inp.clone()
        _, _, self.low, self.high = td_gate_indices(inp.time, self.low, self.high)
        mult = td_gate_multiplier(inp.time, self.low, self.high, self.window_width)
        out.data = (out.data * mult.astype(out.data.dtype)).astype(out.data.dtype)
        return out


# --------------------------------------------------------------------------------------
# TiltCompensation (src/filters/tilt_compensation.rs:97-226)
# --------------------------------------------------------------------------------------
def _linspace32(a, b, n):
    """ndarray `Array1::linspace(a, b, n)` for f32: a + step*i with step=(b-a)/(n-1)."""
    a = F32(a)
    b = F32(b)
    if n == 0:
        return np.zeros(0, F32)
    if n == 1:
        return np.array([a], F32)
    step = (b - a) / F32(n - 1)
    return (a + step * np.arange(n, dtype=F32)).astype(F32)


@dataclass
class TiltCompensation:
    tilt_x: float = 0.0
    tilt_y: float = 0.0

    def num_steps(self, width, height, dx, dy):
        time_shift_x = F32(self.tilt_x) / F32(180.0) * PI32
        time_shift_y = F32(self.tilt_y) / F32(180.0) * PI32
        center_x = F32(width) / F32(2.0) * F32(dx)
        center_y = F32(height) / F32(2.0) * F32(dy)
        c = 0.299792458
        dt = F32(0.05)
        max_offset_x = F32(float(center_x) * float(abs(time_shift_x)) / c)
        max_offset_y = F32(float(center_y) * float(abs(time_shift_y)) / c)
        extension = (max_offset_x + max_offset_y) / dt
        extension = np.floor(extension) * dt
        return int(np.round(extension / dt)), F32(extension), time_shift_x, time_shift_y

    def filter(self, inp: ScannedImageFilterData) -> ScannedImageFilterData:
        out = inp.clone()
        if inp.dx is None or inp.dy is None:
            return out
        dx, dy = F32(inp.dx), F32(inp.dy)
        width, height, n = inp.data.shape
        c = 0.299792458
        dt = F32(0.05)
        if inp.time.shape[0] == 0:
            return out
        num_steps, extension, tsx, tsy = self.num_steps(width, height, dx, dy)
        first_value, last_value = inp.time[0], inp.time[-1]
        ext_n = n + 2 * num_steps
        front = _linspace32(first_value - extension, first_value - dt, num_steps)
        back = _linspace32(last_value + dt, last_value + extension, num_steps)
        out.time = np.concatenate([front, inp.time.astype(F32), back]).astype(F32)
        ddt = inp.data.dtype
        taper = adapted_blackman_multiplier(inp.time, 0.0, 7.0).astype(ddt)
        new = np.zeros((width, height, ext_n), dtype=ddt)
        for i in range(width):
            x_off = F32(float((F32(i) - F32(width) / F32(2.0)) * dx) * float(tsx) / c)
            for j in range(height):
                y_off = F32(float((F32(j) - F32(height) / F32(2.0)) * dy) * float(tsy) / c)
                delta = x_off + y_off
                delta_steps = int(np.floor(delta / dt))
                insert = max(num_steps + delta_steps, 0)
                raw = inp.data[i, j, :]
                ext = np.zeros(ext_n, dtype=ddt)
                ext[:insert] = raw[0]
                end = min(insert + n, ext_n)
                ext[insert:end] = (raw * taper)[: end - insert]
                new[i, j, :] = ext
        out.frequency = frequency_axis(out.time)
        out.data = new
        return out


# --------------------------------------------------------------------------------------
# PSF model (src/filters/psf.rs) and loader (src/io.rs:190-267)
# --------------------------------------------------------------------------------------
@dataclass
class CubicSplineCoeffs:
    knots: np.ndarray
    values: np.ndarray
    coeff_a: np.ndarray
    coeff_b: np.ndarray
    coeff_c: np.ndarray
    coeff_d: np.ndarray

    def _segment(self, x):
        n = self.knots.shape[0]
        left, right = 0, n - 1
        while right - left > 1:
            mid = (left + right) // 2
            if self.knots[mid] > x:
                right = mid
            else:
                left = mid
        return left

    def _poly(self, i, dx):
        a, b, c, d = self.coeff_a[i], self.coeff_b[i], self.coeff_c[i], self.coeff_d[i]
        return a + b * dx + c * dx * dx + d * dx * dx * dx

    def eval_single(self, x):
        """psf.rs:26-80 (linear extrapolation, clamped >= 1e-6 outside the knots)."""
        x = F32(x)
        n = self.knots.shape[0]
        if n == 0:
            return F32(0.0)
        if x < self.knots[0]:
            dx = x - self.knots[0]
            return max(self.coeff_a[0] + self.coeff_b[0] * dx, F32(1e-6))
        if x > self.knots[n - 1]:
            i = n - 2
            dx_end = self.knots[n - 1] - self.knots[i]
            y_end = self._poly(i, dx_end)
            slope_end = (self.coeff_b[i] + F32(2.0) * self.coeff_c[i] * dx_end
                         + F32(3.0) * self.coeff_d[i] * dx_end * dx_end)
            dx = x - self.knots[n - 1]
            return max(y_end + slope_end * dx, F32(1e-6))
        left = self._segment(x)
        return self._poly(left, x - self.knots[left])

    def eval_single_const_extrap(self, x):
        """psf.rs:83-117."""
        x = F32(x)
        n = self.knots.shape[0]
        if n == 0:
            return F32(0.0)
        if x < self.knots[0]:
            return self.values[0]
        if x > self.knots[n - 1]:
            return self.values[n - 1]
        left = self._segment(x)
        return self._poly(left, x - self.knots[left])


@dataclass
class HybridFit:
    base_a: np.float32
    base_b: np.float32
    correction: CubicSplineCoeffs

    def eval_correction(self, f):
        """psf.rs:134-179."""
        f = F32(f)
        c = self.correction
        n = c.knots.shape[0]
        if n == 0:
            return F32(0.0)
        f_min, f_max = c.knots[0], c.knots[n - 1]
        if f_min <= f <= f_max:
            return c.eval_single(f)
        if f < f_min:
            dx = f - f_min
            max_slope = self.base_a / (f * f)
            safe_slope = min(c.coeff_b[0], max_slope)
            return c.coeff_a[0] + safe_slope * dx
        i = n - 2
        dx_end = c.knots[n - 1] - c.knots[i]
        y_end = c._poly(i, dx_end)
        slope_end = (c.coeff_b[i] + F32(2.0) * c.coeff_c[i] * dx_end
                     + F32(3.0) * c.coeff_d[i] * dx_end * dx_end)
        max_slope = self.base_a / (f * f)
        safe_slope = min(slope_end, max_slope)
        dx = f - c.knots[n - 1]
        return y_end + safe_slope * dx

    def eval_single(self, f):
        """psf.rs:122-131."""
        f = F32(f)
        base = self.base_a / f + self.base_b
        return max(F32(base + self.eval_correction(f)), F32(1e-6))


@dataclass
class PSF:
    wx_fit: HybridFit
    wy_fit: HybridFit
    x0_spline: CubicSplineCoeffs
    y0_spline: CubicSplineCoeffs


def load_psf(path) -> PSF:
    """`load_psf` (src/io.rs:190-267): 26 f64 arrays cast to f32."""
    z = np.load(path)

    def arr(name):
        return np.asarray(z[name], dtype=np.float64).reshape(-1).astype(F32)

    def spline(prefix):
        return CubicSplineCoeffs(arr(f"{prefix}_knots_thz"), arr(f"{prefix}_values_mm"),
                                 arr(f"{prefix}_coeff_a"), arr(f"{prefix}_coeff_b"),
                                 arr(f"{prefix}_coeff_c"), arr(f"{prefix}_coeff_d"))

    def hybrid(prefix):
        return HybridFit(arr(f"{prefix}_base_a")[0], arr(f"{prefix}_base_b")[0], spline(f"{prefix}_corr"))

    return PSF(hybrid("wx"), hybrid("wy"), spline("x0"), spline("y0"))


def gaussian(x, params):
    """psf.rs:326-332 (f32; powf(2.0) == square)."""
    x = np.asarray(x, dtype=F32)
    x0, w = F32(params[0]), F32(params[1])
    d = x - x0
    return (np.sqrt(F32(2.0) / PI32) * _exp32(F32(-2.0) * (d * d) / (w * w)) / w).astype(F32)


def create_psf_2d(psf_x, psf_y, x, y, dx, dy):
    """psf.rs:228-313.  The interpolator is only ever queried at its own knots (the
    grid is `v*dx`, the knots are `i*dx`), so it reduces to a lookup; values outside
    the original support are the zero padding."""
    px = np.asarray(psf_x, dtype=F32).copy()
    py = np.asarray(psf_y, dtype=F32).copy()
    x = np.asarray(x, dtype=F32)
    y = np.asarray(y, dtype=F32)
    dx, dy = F32(dx), F32(dy)
    px = px / px.max()
    py = py / py.max()
    x_max = int(np.floor(x.max()))
    y_max = int(np.floor(y.max()))
    kx = (x.shape[0] - 1) // 2
    ky = (y.shape[0] - 1) // 2

    def axis_vals(p, k, vmax):
        out = np.zeros(2 * vmax + 1, dtype=F32)
        for idx, v in enumerate(range(-vmax, vmax + 1)):
            if -k <= v <= k:
                out[idx] = p[v + k]
        return out

    gx = axis_vals(px, kx, x_max)
    gy = axis_vals(py, ky, y_max)
    return np.outer(gx, gy).astype(F32), gx, gy


# --------------------------------------------------------------------------------------
# FIR bank design (src/filters/deconvolution.rs:30-211) -- float64 host arithmetic
# --------------------------------------------------------------------------------------
def kaiser_atten(ntaps, width_ratio):
    return max(2.285 * (float(ntaps) - 1.0) * math.pi * width_ratio + 7.95, 0.0)


def kaiser_beta(atten):
    if atten > 50.0:
        return 0.1102 * (atten - 8.7)
    if atten >= 21.0:
        return 0.5842 * (atten - 21.0) ** 0.4 + 0.07886 * (atten - 21.0)
    return 0.0


def i0(x):
    s = 1.0
    term = 1.0
    x_half_sq = (x / 2.0) ** 2
    for k in range(1, 50):
        term *= x_half_sq / float(k * k)
        s += term
        if term < 1e-12 * s:
            break
    return s


def sinc(x):
    return 1.0 if abs(x) < 1e-10 else math.sin(x) / x


def kaiser_window_coeff(n, n_taps, beta):
    if n == 0 or n == n_taps - 1:
        return 0.0
    arg = 2.0 * n / (float(n_taps) - 1.0) - 1.0
    return i0(beta * math.sqrt(1.0 - arg * arg)) / i0(beta)


def firwin_kaiser_lowpass(n_taps, cutoff_hz, beta, fs):
    adj = n_taps - 1 if n_taps % 2 == 0 else n_taps
    mid = (adj - 1) / 2.0
    cutoff = cutoff_hz / fs
    h = [sinc(2.0 * math.pi * cutoff * (n - mid)) * kaiser_window_coeff(n, adj, beta) for n in range(adj)]
    s = 0.0
    for v in h:
        s += v
    if abs(s) > 1e-10:
        h = [v / s for v in h]
    if n_taps % 2 == 0:
        h.append(0.0)
    return h


def firwin_kaiser_highpass(n_taps, cutoff_hz, beta, fs):
    adj = n_taps - 1 if n_taps % 2 == 0 else n_taps
    mid = (adj - 1) / 2.0
    h = firwin_kaiser_lowpass(adj, cutoff_hz, beta, fs)
    h = [(1.0 - v) if i == int(mid) else -v for i, v in enumerate(h)]
    if n_taps % 2 == 0:
        h.append(0.0)
    return h


def bandpass_kaiser(ntaps, lowcut, highcut, fs, width):
    beta = kaiser_beta(kaiser_atten(ntaps, width / (0.5 * fs)))
    if lowcut <= 0.0:
        return firwin_kaiser_lowpass(ntaps, highcut, beta, fs)
    if highcut >= 0.5 * fs:
        return firwin_kaiser_highpass(ntaps, lowcut, beta, fs)
    lo = firwin_kaiser_highpass(ntaps, lowcut, beta, fs)
    hi = firwin_kaiser_highpass(ntaps, highcut, beta, fs)
    return [a - b for a, b in zip(lo, hi)]


NTAPS = 499


def create_filter_bank(n_filters, start_freq, end_freq, win_width, time_array):
    """deconvolution.rs:160-211 -> (filters[B][499] f32, center_frequencies[B] f32)."""
    t = np.asarray(time_array, dtype=F32)
    dt = float(t[1] - t[0])
    fs = 1.0 / dt
    log_start, log_end = math.log(start_freq), math.log(end_freq)
    log_step = (log_end - log_start) / float(n_filters - 1)
    centers = np.array([F32(math.exp(log_start + i * log_step)) for i in range(n_filters)], dtype=F32)
    filters = np.zeros((n_filters, NTAPS), dtype=F32)
    for i in range(n_filters):
        cf = float(centers[i])
        lowcut = 0.0 if i == 0 else math.sqrt(float(centers[i - 1]) * cf)
        highcut = 0.5 * fs if i == n_filters - 1 else math.sqrt(cf * float(centers[i + 1]))
        filters[i, :] = np.asarray(bandpass_kaiser(NTAPS, lowcut, highcut, fs, win_width), dtype=np.float64).astype(F32)
    return filters, centers


# --------------------------------------------------------------------------------------
# convolve1d / filter_scan (deconvolution.rs:266-317, 574-609)
# --------------------------------------------------------------------------------------
def next_pow2(n):
    return 1 << (int(n) - 1).bit_length() if n > 1 else 1


def filter_scan(data, fir, chunk_rows=None, workers=None):
    """Every trace convolved with the FIR via complex f64 FFTs of next_pow2(N+len-1),
    'same' alignment with shift (len-1)/2, result cast to f32."""
    data = np.asarray(data)
    w, h, n = data.shape
    taps = fir.shape[0]
    m = next_pow2(n + taps - 1)
    shift = (taps - 1) // 2
    hb = np.zeros(m, dtype=np.complex128)
    hb[:taps] = fir.astype(np.float64)
    hf = sfft.fft(hb)
    out = np.zeros((w, h, n), dtype=data.dtype)
    if chunk_rows is None:
        chunk_rows = max(1, int(2 ** 26 // max(1, h * m)))
    for r0 in range(0, w, chunk_rows):
        r1 = min(w, r0 + chunk_rows)
        a = np.zeros((r1 - r0, h, m), dtype=np.complex128)
        a[:, :, :n] = data[r0:r1].astype(np.float64)
        af = sfft.fft(a, axis=-1, workers=workers)
        af *= hf
        y = sfft.ifft(af, axis=-1, norm="forward", workers=workers)
        out[r0:r1] = (y[:, :, shift:shift + n].real / float(m)).astype(data.dtype)
    return out


# --------------------------------------------------------------------------------------
# 2-D "same" filtering with zero boundary (deconvolution.rs:350-545)
# --------------------------------------------------------------------------------------
def direct_convolve2d(a, b):
    """deconvolution.rs:432-458 -- a *correlation*; accumulation order m then n."""
    a = np.asarray(a)
    b = np.asarray(b)
    ar, ac = a.shape
    br, bc = b.shape
    hr, hc = br // 2, bc // 2
    pad = np.zeros((ar + br, ac + bc), dtype=a.dtype)
    pad[hr:hr + ar, hc:hc + ac] = a
    res = np.zeros((ar, ac), dtype=a.dtype)
    for m in range(br):
        for n in range(bc):
            res = res + pad[m:m + ar, n:n + ac] * b[m, n]
    return res


def fft_convolve2d(a, b):
    """The FFT branch of `convolve2d` (deconvolution.rs:489-544): complex FFTs in the
    image precision at next_pow2(dim+k-1), a true convolution, crop at (k-1)/2."""
    a = np.asarray(a)
    b = np.asarray(b)
    ar, ac = a.shape
    br, bc = b.shape
    pr = next_pow2(ar + br - 1)
    pc = next_pow2(ac + bc - 1)
    cdt = _complex_of(a.dtype)
    ap = np.zeros((pr, pc), dtype=cdt)
    bp = np.zeros((pr, pc), dtype=cdt)
    ap[:ar, :ac] = a
    bp[:br, :bc] = b
    af = sfft.fft2(ap)
    bf = sfft.fft2(bp)
    res = sfft.ifft2((af * bf).astype(cdt), norm="forward")
    res = (res / a.dtype.type(pr * pc)).astype(cdt)
    sr, sc = (br - 1) // 2, (bc - 1) // 2
    return np.ascontiguousarray(res[sr:sr + ar, sc:sc + ac].real).astype(a.dtype)


CONV2D_THRESHOLD = 256


def convolve2d(a, b):
    """deconvolution.rs:472-545."""
    if b.shape[0] * b.shape[1] <= CONV2D_THRESHOLD:
        return direct_convolve2d(a, b)
    return fft_convolve2d(a, b)


def reflect_pad(image, pad_y, pad_x):
    """The hand-written numpy-'reflect' padding of `richardson_lucy`
    (deconvolution.rs:629-667); axis 0 is padded by pad_y = psf.nrows()/2."""
    h, w = image.shape
    p = np.zeros((h + 2 * pad_y, w + 2 * pad_x), dtype=image.dtype)
    p[pad_y:pad_y + h, pad_x:pad_x + w] = image
    for i in range(pad_y):
        p[i, pad_x:pad_x + w] = image[pad_y - i, :]
        p[pad_y + h + i, pad_x:pad_x + w] = image[h - 2 - i, :]
    for j in range(pad_x):
        p[:, j] = p[:, pad_x + (pad_x - j)].copy()
        p[:, pad_x + w + j] = p[:, pad_x + w - 2 - j].copy()
    return p


def richardson_lucy(image, psf, n_iterations, conv=convolve2d, return_padded=False):
    """deconvolution.rs:620-712."""
    image = np.asarray(image)
    psf = np.asarray(psf, dtype=image.dtype)
    psf_mirror = np.ascontiguousarray(psf[::-1, ::-1])
    pad_y, pad_x = psf.shape[0] // 2, psf.shape[1] // 2
    h, w = image.shape
    padded = reflect_pad(image, pad_y, pad_x)
    u = padded.copy()
    eps = image.dtype.type(1e-12)
    for _ in range(int(n_iterations)):
        ustarp = conv(u, psf)
        relative_blur = padded / (ustarp + eps)
        correction = conv(relative_blur, psf_mirror)
        u = u * correction
    if return_padded:
        return u
    return np.ascontiguousarray(u[pad_y:pad_y + h, pad_x:pad_x + w])


# --------------------------------------------------------------------------------------
# Deconvolution::filter orchestration (deconvolution.rs:766-1041)
# --------------------------------------------------------------------------------------
@dataclass
class BandPlan:
    """Everything `Deconvolution::filter` derives per band before touching the cube."""
    center_freq: np.float32
    fir: np.ndarray          # [499] f32
    wx: np.float32
    wy: np.float32
    x0: np.float32
    y0: np.float32
    psf: np.ndarray          # [kx][ky] f32 (axis 0 <-> x)
    psf_x: np.ndarray        # separable factors: psf = outer(psf_x, psf_y)
    psf_y: np.ndarray
    n_iter: int


@dataclass
class Deconvolution:
    n_iterations: int = 500
    n_filters: int = 25
    start_freq: float = 0.1
    end_freq: float = 10.0
    win_width: float = 0.5

    MIN_IMAGE_SIZE = 16

    def plan(self, time, shape, dx, dy, psf: PSF):
        """Returns (list[BandPlan], None) or (None, reason) when the reference skips the
        filter and returns its input (deconvolution.rs:781-885)."""
        if dx is None or dy is None:
            return None, "no dx/dy"
        if psf is None or psf.wx_fit.correction.knots.shape[0] == 0:
            return None, "no psf"
        img_rows, img_cols = int(shape[0]), int(shape[1])
        if img_rows < self.MIN_IMAGE_SIZE or img_cols < self.MIN_IMAGE_SIZE:
            return None, "image too small"
        # start_freq / end_freq / win_width are f32 fields cast to f64 (:821-826)
        filters, centers = create_filter_bank(self.n_filters, float(F32(self.start_freq)),
                                              float(F32(self.end_freq)), float(F32(self.win_width)), time)
        wx_values = np.array([psf.wx_fit.eval_single(f) for f in centers], dtype=F32)
        wy_values = np.array([psf.wy_fit.eval_single(f) for f in centers], dtype=F32)
        w_min = min(wx_values.min(), wy_values.min())
        w_max = max(wx_values.max(), wy_values.max())
        dx, dy = F32(dx), F32(dy)
        max_psf_width_x = max(int(np.ceil(wx_values.max() / dx)) * 2 + 1, 3)
        max_psf_width_y = max(int(np.ceil(wy_values.max() / dy)) * 2 + 1, 3)
        if max_psf_width_x >= img_cols or max_psf_width_y >= img_rows:
            return None, "psf too large"
        bands = []
        for i in range(self.n_filters):
            cf = centers[i]
            wx = psf.wx_fit.eval_single(cf)
            wy = psf.wy_fit.eval_single(cf)
            x0 = psf.x0_spline.eval_single_const_extrap(cf)
            y0 = psf.y0_spline.eval_single_const_extrap(cf)
            rx = max(F32((wx + abs(x0)) * F32(3.0)), F32(2.5))
            ry = max(F32((wy + abs(y0)) * F32(3.0)), F32(2.5))
            rx = np.floor(rx / dx) * dx + dx
            ry = np.floor(ry / dy) * dy + dy
            max_allowed_x = (F32(img_cols) - F32(2.0)) * dx / F32(2.0)
            max_allowed_y = (F32(img_rows) - F32(2.0)) * dy / F32(2.0)
            crx = min(rx, max_allowed_x)
            cry = min(ry, max_allowed_y)
            kx = int(np.floor(crx / dx))
            ky = int(np.floor(cry / dy))
            x = (np.arange(-kx, kx + 1, dtype=F32) * dx).astype(F32)
            y = (np.arange(-ky, ky + 1, dtype=F32) * dy).astype(F32)
            gx = gaussian(x, (x0, wx))
            gy = gaussian(y, (y0, wy))
            psf2d, px, py = create_psf_2d(gx, gy, x, y, dx, dy)
            n_iter = int(np.floor((wx - w_min) / (w_max - w_min) * (F32(self.n_iterations) - F32(1.0)) + F32(1.0)))
            bands.append(BandPlan(cf, filters[i].copy(), wx, wy, x0, y0, psf2d, px, py, n_iter))
        return bands, None

    def filter(self, inp: ScannedImageFilterData, psf: PSF, conv=convolve2d, workers=None,
               return_intermediates=False):
        bands, reason = self.plan(inp.time, inp.data.shape, inp.dx, inp.dy, psf)
        if bands is None:
            return (inp.clone(), None) if return_intermediates else inp.clone()
        out = inp.clone()
        acc = np.zeros_like(inp.data)
        inter = []
        for b in bands:
            filtered = filter_scan(inp.data, b.fir, workers=workers)
            filtered_image = np.sum(filtered * filtered, axis=2, dtype=filtered.dtype)
            u = richardson_lucy(filtered_image, b.psf, b.n_iter, conv=conv)
            u = np.maximum(u, u.dtype.type(0.0))
            with np.errstate(divide="ignore", invalid="ignore"):
                gains = np.sqrt(u / filtered_image)
            acc = acc + filtered * gains[:, :, None]
            inter.append((filtered_image, u, gains))
        out.data = acc
        out.img = intensity_image(acc)
        return (out, inter) if return_intermediates else out


# --------------------------------------------------------------------------------------
# the default chain (src/main.rs:194-247 order; data_thread.rs:1090-1228)
# --------------------------------------------------------------------------------------
@dataclass
class ChainParams:
    """Filter parameters of the default chain, reference defaults."""
    config: ConfigContainer = field(default_factory=ConfigContainer)
    tilt: TiltCompensation = field(default_factory=TiltCompensation)
    gate_before: TimeDomainBandPass = field(default_factory=lambda: TimeDomainBandPass(window_width=2.0))
    band: FrequencyDomainBandPass = field(default_factory=FrequencyDomainBandPass)
    gate_after: TimeDomainBandPass = field(default_factory=lambda: TimeDomainBandPass(window_width=0.1))
    tilt_active: bool = True
    gate_before_active: bool = True
    band_active: bool = True
    gate_after_active: bool = True


def run_default_chain(slot0: ScannedImageFilterData, params: ChainParams = None, workers=None,
                      reset=True):
    """Slots 1..7 of the default chain: scaling -> tilt -> gate -> fft -> band-pass ->
    ifft -> gate.  Returns the list of pipeline slots (slot 8, deconvolution, is run
    separately because it only fires on its Apply button, data_thread.rs:1139-1150)."""
    p = params or ChainParams()
    if reset:
        # filter.reset(time, shape) for every filter when a file was opened
        # (data_thread.rs:1027-1060); each gets the time axis of its input slot, which
        # at 0 deg tilt is the scan's own axis.
        p.gate_before.reset(slot0.time)
        p.gate_after.reset(slot0.time)
    slots = [slot0]
    slots.append(scaling(slots[-1], p.config))
    slots.append(p.tilt.filter(slots[-1]) if p.tilt_active else slots[-1].clone())
    slots.append(p.gate_before.filter(slots[-1]) if p.gate_before_active else slots[-1].clone())
    slots.append(fft(slots[-1], p.config, workers=workers))
    slots.append(p.band.filter(slots[-1]) if p.band_active else slots[-1].clone())
    slots.append(ifft(slots[-1], p.config, workers=workers))
    slots.append(p.gate_after.filter(slots[-1]) if p.gate_after_active else slots[-1].clone())
    slots[-1].img = intensity_image(slots[-1].data)
    return slots


def default_chain_multipliers(time, dx_dy_present=True, params: ChainParams = None):
    """The three host-side vectors the default chain reduces to:
    m_pre[N]  = tilt taper * gate-before * fft window   (sequential f32 products)
    band[F]   = FD band-pass multiplier
    m_post[N] = gate-after.
    Only valid at 0 deg tilt (no time-axis extension)."""
    p = params or ChainParams()
    t = np.asarray(time, dtype=F32)
    m_pre = np.ones(t.shape[0], dtype=F32)
    if p.tilt_active and dx_dy_present:
        m_pre = m_pre * adapted_blackman_multiplier(t, 0.0, 7.0)
    if p.gate_before_active:
        gb = TimeDomainBandPass(window_width=p.gate_before.window_width)
        gb.reset(t)
        m_pre = m_pre * td_gate_multiplier(t, gb.low, gb.high, gb.window_width)
    m_pre = m_pre * fft_window_multiplier(t, p.config.fft_window_type, p.config.fft_window)
    f = frequency_axis(t)
    band = (fd_band_multiplier(f, p.band.low, p.band.high, p.band.window_width)
            if p.band_active else np.ones(f.shape[0], F32))
    if p.gate_after_active:
        ga = TimeDomainBandPass(window_width=p.gate_after.window_width)
        ga.reset(t)
        m_post = td_gate_multiplier(t, ga.low, ga.high, ga.window_width)
    else:
        m_post = np.ones(t.shape[0], F32)
    return m_pre.astype(F32), band.astype(F32), m_post.astype(F32)


# --------------------------------------------------------------------------------------
# optical properties (src/math_tools.rs:663-701) -- "next" row, restated for completeness
# --------------------------------------------------------------------------------------
C_LIGHT = F32(2.99792458e8)


def calculate_optical_properties(sample_amp, sample_phase, ref_amp, ref_phase, freqs, thickness):
    sa, sp = np.asarray(sample_amp, F32), np.asarray(sample_phase, F32)
    ra, rp = np.asarray(ref_amp, F32), np.asarray(ref_phase, F32)
    f = np.asarray(freqs, F32)
    d = F32(thickness)
    with np.errstate(divide="ignore", invalid="ignore"):
        f_hz = f * F32(1.0e12)
        dphi = sp - rp
        omega = F32(2.0) * PI32 * f_hz
        n = F32(1.0) + C_LIGHT * dphi / (omega * d)
        amp = np.maximum(sa, F32(1e-12))
        amp_ref = np.maximum(ra, F32(1e-12))
        n_safe = np.maximum(n, F32(1e-6))
        arg = ((n_safe + F32(1.0)) ** 2) / (F32(4.0) * n_safe) * amp / amp_ref
        alpha = F32(-2.0) / d * np.log(arg.astype(np.float64)).astype(F32)
        kappa = alpha * C_LIGHT / (F32(4.0) * PI32 * f_hz)
    return n.astype(F32), alpha.astype(F32), kappa.astype(F32)


# --------------------------------------------------------------------------------------
# ROI averages (src/math_tools.rs:572-661) -- "next" row, restated for completeness
# --------------------------------------------------------------------------------------
_U64 = (1 << 64) - 1


def point_in_polygon(x, y, polygon):
    """Ray casting with the reference's `usize` arithmetic (release build: wrapping)."""
    inside = False
    j = len(polygon) - 1
    for i in range(len(polygon)):
        xi, yi = polygon[i]
        xj, yj = polygon[j]
        if (yi > y) != (yj > y):
            num = (((xj - xi) & _U64) * ((y - yi) & _U64)) & _U64
            den = (yj - yi) & _U64
            if x < (((num // den) + xi) & _U64):
                inside = not inside
        j = i
    return inside


def average_polygon_roi(data, polygon, scaling=1):
    """`average_polygon_roi` (src/math_tools.rs:599-661): sequential f32 sums over the pixels inside the
    polygon (bounding box scan y outer, x inner; row index flipped: data[y_size - y - 1, x, :])."""
    data = np.asarray(data, dtype=F32)
    poly = [(int(x) // scaling, int(y) // scaling) for x, y in polygon]
    y_size, x_size, z_size = data.shape
    x_min = min(min(p[0] for p in poly), x_size - 1)
    y_min = min(min(p[1] for p in poly), y_size - 1)
    x_max = min(max(p[0] for p in poly), x_size - 1)
    y_max = min(max(p[1] for p in poly), y_size - 1)
    result = np.zeros(z_size, dtype=F32)
    count = 0
    for y in range(y_min, y_max + 1):
        for x in range(x_min, x_max + 1):
            if point_in_polygon(x, y, poly):
                result = result + data[y_size - y - 1, x, :]
                count += 1
    if count > 0:
        result = result / F32(count)
    return result.astype(F32)


# --------------------------------------------------------------------------------------
# reference pulse, ConfigCommand::OpenRef (src/data_thread.rs:372-588) -- "next" row
# --------------------------------------------------------------------------------------
def reference_pulse(scan_time, ref_time, ref_signal, config: ConfigContainer = None):
    """Aligns / resizes a reference pulse to the scan's time axis (integer shift, zero fill,
    data_thread.rs:405-481), windows it on the REFERENCE file's time axis (:489-511, the zip stops at the
    shorter of the two), then r2c + |s| + unwrap(arg s) with the scan's plan (:513-533).
    Returns (signal[n], amplitudes[F], phases[F])."""
    config = config or ConfigContainer()
    st = np.asarray(scan_time, F32)
    rt = np.asarray(ref_time, F32)
    ref = np.asarray(ref_signal, F32).copy()
    n = st.shape[0]
    if n != ref.shape[0] or (rt.shape[0] and abs(float(st[0] - rt[0])) > 1e-9):
        if n > 1 and rt.shape[0] > 1:
            new = np.zeros(n, F32)
            ref_dt = rt[1] - rt[0]
            time_offset = st[0] - rt[0]
            index_offset = int(np.round(time_offset / ref_dt))   # f32 division, round half away from zero
            if abs(float(time_offset / ref_dt)) % 1 == 0.5:
                index_offset = int(np.sign(float(time_offset / ref_dt)) * np.ceil(abs(float(time_offset / ref_dt))))
            src_start = index_offset if index_offset > 0 else 0
            dst_start = -index_offset if index_offset < 0 else 0
            copy_len = min(max(ref.shape[0] - src_start, 0), max(n - dst_start, 0))
            if copy_len > 0:
                new[dst_start:dst_start + copy_len] = ref[src_start:src_start + copy_len]
            ref = new
        else:
            new = np.zeros(n, F32)
            k = min(n, ref.shape[0])
            new[:k] = ref[:k]
            ref = new
    mult = fft_window_multiplier(rt, config.fft_window_type, config.fft_window)
    k = min(n, rt.shape[0])
    ref[:k] = ref[:k] * mult[:k]
    spec = rfft_unnormalised(ref).astype(np.complex64)
    amp = np.abs(spec).astype(F32)
    ph = numpy_unwrap(np.arctan2(spec.imag, spec.real).astype(F32), F32(2.0) * PI32).astype(F32)
    return ref, amp, ph


# --------------------------------------------------------------------------------------
# voxel envelope of the 3-D view (src/gui/threed_plot.rs:81-236) -- "next" row
# --------------------------------------------------------------------------------------
def gaussian_kernel1d(sigma, radius):
    sigma = F32(sigma)
    x = np.arange(2 * radius + 1, dtype=F32) - F32(radius)
    v = _exp32(-x * x / (F32(2.0) * sigma * sigma))
    s = F32(0.0)
    for t in v:
        s = F32(s + t)
    return (v / s).astype(F32)


def voxel_opacity(dataset, opacity_threshold=0.1, contrast=2.0, sigma=3.0, radius=9, max_instances=2_000_000):
    """`instance_from_data` up to the effective threshold (threed_plot.rs:165-219): per trace v^2,
    Gaussian-weighted sum of (v^2)^contrast with zero boundary (taps accumulated in order), traces whose
    maximum is below opacity_threshold are zeroed, the others min/max normalised; the effective threshold is
    the max_instances-th largest opacity (0 when the cube has fewer voxels).  Returns (opacity, threshold)."""
    d = np.asarray(dataset, F32)
    kernel = gaussian_kernel1d(sigma, radius)
    sq = (d * d).astype(F32)
    pw = np.power(sq.astype(np.float64), float(F32(contrast))).astype(F32)
    n = d.shape[-1]
    env = np.zeros_like(d)
    for k, coeff in enumerate(kernel):
        sh = k - radius
        lo, hi = max(0, -sh), min(n, n - sh)
        if hi > lo:
            env[..., lo:hi] = env[..., lo:hi] + pw[..., lo + sh:hi + sh] * coeff
    mx = env.max(axis=-1, keepdims=True)
    mn = env.min(axis=-1, keepdims=True)
    rng = mx - mn
    ok = (mx >= F32(opacity_threshold)) & (np.abs(rng) > F32(1e-6))
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(ok, (env - mn) / rng, F32(0.0)).astype(F32)
    flat = out.reshape(-1)
    thr = F32(0.0)
    if flat.size > max_instances:
        thr = np.partition(flat, flat.size - max_instances)[flat.size - max_instances]
    return out, F32(thr)
