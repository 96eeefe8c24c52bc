"""The algebra behind the round-2 restructurings of the cube passes (DESIGN.md section 2-11 ... 2-13), checked in
numpy float64 against straightforward transforms.  No GPU, no library call: these pin the identities the kernels
rely on, so that a GPU parity failure can be told apart from a wrong derivation.

  * even bins of the 2N-point spectrum of the gated trace from the filtered spectrum plus an R-point DFT of the
    gate's edge samples (k_chain_energy_fused, POST = 1);
  * odd bins as the N-point transform of the modulated trace, Parseval energy of the full linear FIR convolution
    from the two sub-spectra of a pair-packed transform (q1 = p + c, q2 = p - c);
  * the spectral hand-off: pass C from FFT_N(y) equals pass C from y (circular form);
  * the chirp-z convolution on a 2M-point frame as two M-point sub-spectra (k_blue_*<M, SPLIT>)."""
import numpy as np
import pytest


def _gate(n, rng):
    m = np.ones(n)
    m[:4] = [0.0, 0.34, 0.83, 0.999]
    m[-4:] = rng.uniform(0.0, 1.0, 4)
    return m


@pytest.mark.parametrize("n,radix", [(4096, 16), (2048, 8), (512, 4), (1024, 4)])
def test_even_bins_from_filtered_spectrum_and_sparse_gate(n, radix):
    """FFT_N(m * IFFT(Z'))[k0 + q N/R] = N Z'[k] - DFT_R(d)[q],  d[s] = sum over n == s (mod R) of c_n w_N^(k0 n),
    c_n = y0[n] - y[n] non-zero only in the first / last four samples."""
    rng = np.random.default_rng(n)
    zp = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / n        # Z' = (band / N) X of a packed pair
    y0 = np.fft.ifft(zp) * n                                               # unnormalised inverse, as the kernel runs it
    m = _gate(n, rng)
    y = m * y0
    direct = np.fft.fft(y)
    c = y0 - y
    support = np.nonzero(c)[0]
    assert np.all((support < 4) | (support >= n - 4))
    got = np.empty(n, complex)
    for k0 in range(n // radix):                                           # one (thread, u) of the kernel
        w1 = np.exp(-2j * np.pi * k0 / n)
        d = np.zeros(radix, complex)
        for j in range(4):
            d[j % radix] += c[j] * w1 ** j                                 # head samples n = j
            d[(radix - 1 - j) % radix] += c[n - 1 - j] * np.conj(w1) ** (j + 1)   # tail samples n = N - (j + 1)
        corr = np.fft.fft(d)                                               # forward R-point DFT, natural order
        for q in range(radix):
            k = k0 + q * (n // radix)
            got[k] = n * zp[k] - corr[q]
    assert np.max(np.abs(got - direct)) <= 1e-10 * np.max(np.abs(direct))


def test_band_energy_of_the_linear_convolution_from_two_sub_spectra():
    """sum_t (h * y)[t]^2 over the FULL linear convolution = (1/M) sum_f |H[f]|^2 |Y[f]|^2 on the M = 2N frame, with
    Y[2j] = FFT_N(y)[j], Y[2j+1] = FFT_N(y w_M^n)[j]; for a pair packed as y1 + i y2 the two traces separate as
    |Y1|^2 + |Y1 mirror|^2 = 2 (p + c), |Y2|^2 + ... = 2 (p - c) with p = (|Z|^2 + |Zm|^2) / 2, c = Re(Z Zm)."""
    rng = np.random.default_rng(1)
    n, taps = 512, 499
    m2 = 2 * n
    h = rng.standard_normal(taps)
    h = h + h[::-1]                                                        # symmetric (zero phase once centred)
    y1, y2 = rng.standard_normal(n), rng.standard_normal(n)
    want = [np.sum(np.convolve(h, y) ** 2) for y in (y1, y2)]
    H2 = np.abs(np.fft.fft(h, m2)) ** 2
    z = y1 + 1j * y2
    mod = np.exp(-2j * np.pi * np.arange(n) / m2)
    sub = {0: np.fft.fft(z), 1: np.fft.fft(z * mod)}
    e1 = e2 = 0.0
    for odd, Z in sub.items():
        for j in range(n):
            jm = (n - 1 - j) if odd else (n - j) % n                       # mirror stays inside the sub-spectrum
            zz, zm = Z[j], Z[jm]
            p = 0.5 * (abs(zz) ** 2 + abs(zm) ** 2)
            c = zz.real * zm.real - zz.imag * zm.imag
            w = H2[2 * j + odd] / m2
            e1 += 0.5 * w * (p + c)                                        # every (j, mirror) pair is visited twice
            e2 += 0.5 * w * (p - c)
    assert abs(e1 - want[0]) <= 1e-9 * want[0] and abs(e2 - want[1]) <= 1e-9 * want[1]


def test_pass_c_from_the_spectrum_equals_pass_c_from_the_trace():
    """Circular pass C with per-trace gains on a packed pair: IFFT(S Z + D conj(Z mirror)) = g1 (h (*) y1) + i g2 (h (*) y2),
    S = (g1 + g2) H / 2, D = (g1 - g2) H / 2 -- whether Z was recomputed from the stored traces or handed over."""
    rng = np.random.default_rng(2)
    n = 1024
    hs = np.zeros(n)
    taps = rng.standard_normal(250)
    hs[:250] = taps
    hs[-249:] = taps[1:][::-1]                                             # real, even -> real, even spectrum
    H = np.fft.fft(hs).real
    y1, y2 = rng.standard_normal(n), rng.standard_normal(n)
    g1, g2 = 0.7, 1.9
    Z = np.fft.fft(y1 + 1j * y2)                                           # what the fused kernel hands over
    Zm = np.conj(Z[(-np.arange(n)) % n])
    out = np.fft.ifft(0.5 * (g1 + g2) * H * Z + 0.5 * (g1 - g2) * H * Zm)
    ref1 = g1 * np.fft.ifft(np.fft.fft(y1) * H).real
    ref2 = g2 * np.fft.ifft(np.fft.fft(y2) * H).real
    assert np.max(np.abs(out.real - ref1)) <= 1e-10 * np.max(np.abs(ref1))
    assert np.max(np.abs(out.imag - ref2)) <= 1e-10 * np.max(np.abs(ref2))


@pytest.mark.parametrize("n", [37, 50, 64])
def test_chirp_z_on_a_split_frame(n):
    """Bluestein's DFT of length n with the 2M-point chirp convolution evaluated as two M-point sub-spectra
    (n <= M, the sequence fills at most half of the frame)."""
    rng = np.random.default_rng(n)
    M = 64
    m2 = 2 * M
    assert 2 * n - 1 <= m2 and n <= M
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    k = np.arange(n)
    chirp = np.exp(-1j * np.pi * ((k * k) % (2 * n)) / n)
    b = np.zeros(m2, complex)
    b[:n] = np.conj(chirp)
    b[m2 - n + 1:] = np.conj(chirp[1:][::-1])
    bhat = np.fft.fft(b) / m2
    a = np.zeros(M, complex)
    a[:n] = x * chirp
    v = np.exp(-2j * np.pi * np.arange(M) / m2)
    a_even, a_odd = np.fft.fft(a), np.fft.fft(a * v)                       # A[2j], A[2j+1]
    c_even = np.fft.ifft(a_even * bhat[0::2]) * M                          # unnormalised inverses
    c_odd = np.fft.ifft(a_odd * bhat[1::2]) * M
    c = c_even + np.conj(v) * c_odd
    got = chirp * c[:n]
    want = np.fft.fft(x)
    assert np.max(np.abs(got - want)) <= 1e-10 * np.max(np.abs(want))


def test_same_window_energy_and_circular_pass_c_from_edge_pieces():
    """with the centred 499-tap FIR the reference keeps samples [249, 249 + N) of the full linear
    convolution.  (a) Its energy = energy of the full convolution - |head|^2 - |tail|^2 with head = T x[0..249),
    tail = T' x[N-249..N) two fixed triangular Toeplitz products (the tensor-core GEMM).  (b) The N-point circular
    convolution equals the kept samples plus the same head / tail pieces wrapped around (the edge corrections)."""
    rng = np.random.default_rng(3)
    n, taps, half = 1024, 499, 249
    h = rng.standard_normal(taps)
    x = rng.standard_normal(n)
    full = np.convolve(h, x)                                               # length n + 498
    same = full[half:half + n]
    head, tail = full[:half], full[half + n:]
    # (a) the cut-off pieces as triangular Toeplitz products of the first / last 249 samples
    T_head = np.array([[h[k - j] if k >= j else 0.0 for j in range(half)] for k in range(half)])
    T_tail = np.array([[h[half + 1 + k + (half - 1 - j)] if k + (half - 1 - j) + half + 1 < taps else 0.0
                        for j in range(half)] for k in range(half)])
    assert np.allclose(T_head @ x[:half], head, rtol=0, atol=1e-10 * np.abs(full).max())
    assert np.allclose(T_tail @ x[n - half:], tail, rtol=0, atol=1e-10 * np.abs(full).max())
    e_full = np.sum(np.abs(np.fft.fft(h, 2 * n)) ** 2 * np.abs(np.fft.fft(x, 2 * n)) ** 2) / (2 * n)
    assert abs((e_full - head @ head - tail @ tail) - same @ same) <= 1e-9 * (same @ same)
    # (b) circular convolution with the zero-phase (centred) filter
    hc = np.zeros(n)
    hc[:half + 1] = h[half:]
    hc[-half:] = h[:half]
    circ = np.fft.ifft(np.fft.fft(hc) * np.fft.fft(x)).real
    wrapped = same.copy()
    wrapped[:half] += tail                                                 # pushed past the end -> wraps to the start
    wrapped[n - half:] += head                                             # pushed before the start -> wraps to the end
    assert np.max(np.abs(circ - wrapped)) <= 1e-10 * np.abs(full).max()
