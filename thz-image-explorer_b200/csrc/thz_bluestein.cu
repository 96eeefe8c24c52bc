// thz_bluestein.cu -- the trace pass for trace lengths that are not a power of two.
//
// Real scans have arbitrary N (realfft / rustfft accept any length: src/io.rs:614-628 plans whatever the
// file holds, and TiltCompensation extends the axis by 2 * num_steps, src/filters/tilt_compensation.rs:
// 140-150).  The N-point DFT is evaluated exactly as a chirp-z (Bluestein) convolution on the power-of-two
// machinery of thz_fft.cuh:
//     X[k] = w[k] * sum_n (x[n] w[n]) * conj(w[k - n]),        w[n] = exp(-i pi n^2 / N)
// i.e. a = x w zero-padded to M >= 2N - 1, c = IFFT_M(FFT_M(a) * Bhat), X[k] = w[k] c[k], with
// Bhat = FFT_M(conj w placed circularly) / M precomputed on the host in double precision.  The inverse
// N-point DFT is conj(DFT(conj(.))).  Two real traces are packed as re / im of one complex sequence as
// in the power-of-two kernels; everything stays in registers / shared memory between the HBM read and
// the HBM write.  One CTA of M/16 (>= 256) threads per pair-group, M in [64, 8192] => N <= 4096.
//
// 4096 < N <= 8192 needs M = 16384, for which no register plan exists (1024 threads x 128 registers).  The
// sequence a = x w occupies [0, N) of the M-point frame with N <= M/2, so the M-point transforms split into two
// 8192-point ones exactly like the zero-padded FIR transforms of thz_deconv.cu:
//     A[2j] = FFT_8192(a)[j],   A[2j+1] = FFT_8192(a v)[j],   v[n] = exp(-2 pi i n / M)
//     c[n]  = inv_8192(A_even Bhat_even)[n] + conj(v[n]) inv_8192(A_odd Bhat_odd)[n]      (n < 8192)
// (SPLIT = true: the 8192-point geometry, a thread-private shared-memory stash holds a, then the even half of c).
#include "thz_fft.cuh"
#include "thz_internal.h"

#include <math.h>

namespace thz {

template <int M> struct BGeo {
  static constexpr int T = M / kE;
  static constexpr int NT = (T >= 256) ? T : 256;
  static constexpr int G = NT / T;
  static constexpr int kScr = (32 + kNzWords) * G;
  static constexpr size_t smem_bytes = (size_t)G * padded_len(M) * sizeof(float2) + kScr * sizeof(float);
  static constexpr size_t stash_off = (smem_bytes + 15) & ~(size_t)15;          // SPLIT kernels: [G][M] float2 behind it
  static constexpr size_t smem_bytes_split = stash_off + (size_t)G * M * sizeof(float2);
  static constexpr int kMinBlocks = (NT == 256) ? 2 : 1;
};

struct BlueArgs {
  const float* in;        // [P][n]
  float* out;             // [P][n]
  float* img;
  const float* m_pre;     // [n] or null
  const float* m_post;    // [n] or null
  const float* hn;        // [n] band mirrored to all n bins, / n, natural order
  const float* band;      // [F] or null (inverse kernel)
  const float2* chirp;    // [n] w[k] = exp(-i pi k^2 / n)
  const float2* bhat;     // [M] FFT_M(conj chirp, circular) / M in last-stage register order
                          // (split plans: even bins [8192] | odd bins [8192] of the 16384-point spectrum | v[0..512))
  const float2* tw;
  float2* fft;            // [P][F]
  const float2* fft_in;
  float* amp;
  float* phase;
  float* win;
  int n;
  int64_t P;
};

// multiply by v[n] = exp(-2 pi i n / (2M)) (CONJ: by its conjugate) for n = t + i*T: v[t + i*T] = v[t] exp(-i pi i / 16)
template <int M, bool CONJ>
__device__ __forceinline__ void blue_modulate(float2 (&z)[kE], const float2* __restrict__ mod, int t) {
  static_assert(kE == 16, "rotation constants are exp(-i pi i / 16)");
  constexpr float kC[16] = {1.0f, 0.98078528f, 0.923879533f, 0.831469612f, 0.707106781f, 0.555570233f, 0.382683432f, 0.195090322f, 0.0f, -0.195090322f, -0.382683432f, -0.555570233f, -0.707106781f, -0.831469612f, -0.923879533f, -0.98078528f};
  constexpr float kS[16] = {0.0f, -0.195090322f, -0.382683432f, -0.555570233f, -0.707106781f, -0.831469612f, -0.923879533f, -0.98078528f, -1.0f, -0.98078528f, -0.923879533f, -0.831469612f, -0.707106781f, -0.555570233f, -0.382683432f, -0.195090322f};
  const float2 w0 = __ldg(mod + t);
#pragma unroll
  for (int i = 0; i < kE; ++i) {
    const float2 w = make_float2(w0.x * kC[i] - w0.y * kS[i], w0.x * kS[i] + w0.y * kC[i]);
    z[i] = CONJ ? cmul_conj(z[i], w) : cmul(z[i], w);
  }
}

// N-point DFT of the packed sequence held in stage-0 layout (v[i] <-> element t + i*T, zero for >= n).
// SPLIT: the chirp convolution runs on a 2M-point frame as two M-point sub-spectra (see the file header);
// `stash` is this thread's private column of a [M] float2 shared-memory buffer.
template <int M, bool SPLIT>
__device__ __forceinline__ void bluestein_dft(float2 (&v)[kE], int t, float2* sm, const BlueArgs& a, float2* stash) {
  constexpr int T = BGeo<M>::T;
  constexpr int LAST = Plan<M>::ns - 1;
  constexpr int RL = Plan<M>::r[LAST];
  constexpr int UL = kE / RL;
#pragma unroll
  for (int i = 0; i < kE; ++i) {
    const int e = t + i * T;
    v[i] = (e < a.n) ? cmul(v[i], __ldg(a.chirp + e)) : make_float2(0.f, 0.f);
  }
  if constexpr (SPLIT) {
#pragma unroll
    for (int i = 0; i < kE; ++i) stash[i * T] = v[i];
  }
  fft_forward<M>(v, t, sm, a.tw);
#pragma unroll
  for (int i = 0; i < kE; ++i) {
    const int u = i % UL, m = i / UL;
    v[i] = cmul(v[i], __ldg(a.bhat + m * (M / RL) + t + u * T));
  }
  fft_inverse<M>(v, t, sm, a.tw);
  if constexpr (SPLIT) {
    float2 z[kE];
#pragma unroll
    for (int i = 0; i < kE; ++i) {   // a back into registers, the even half of c into the stash
      z[i] = stash[i * T];
      stash[i * T] = v[i];
    }
    blue_modulate<M, false>(z, a.bhat + 2 * M, t);
    fft_forward<M>(z, t, sm, a.tw);
#pragma unroll
    for (int i = 0; i < kE; ++i) {
      const int u = i % UL, m = i / UL;
      z[i] = cmul(z[i], __ldg(a.bhat + M + m * (M / RL) + t + u * T));
    }
    fft_inverse<M>(z, t, sm, a.tw);
    blue_modulate<M, true>(z, a.bhat + 2 * M, t);
#pragma unroll
    for (int i = 0; i < kE; ++i) v[i] = cadd(stash[i * T], z[i]);
  }
#pragma unroll
  for (int i = 0; i < kE; ++i) {
    const int e = t + i * T;
    v[i] = (e < a.n) ? cmul(v[i], __ldg(a.chirp + e)) : make_float2(0.f, 0.f);
  }
}

template <int M>
__device__ __forceinline__ void blue_load(float2 (&v)[kE], const BlueArgs& a, int t, bool act0, bool act1, int64_t p0,
                                          bool& nz0, bool& nz1) {
  constexpr int T = BGeo<M>::T;
  const float* r0 = a.in + p0 * a.n;
  const float* r1 = r0 + a.n;
  nz0 = nz1 = false;
#pragma unroll
  for (int i = 0; i < kE; ++i) {
    const int e = t + i * T;
    const bool in = e < a.n;
    float x0 = (act0 && in) ? __ldcs(r0 + e) : 0.f, x1 = (act1 && in) ? __ldcs(r1 + e) : 0.f;
    nz0 |= (x0 != 0.f);
    nz1 |= (x1 != 0.f);
    if (a.m_pre != nullptr && in) {
      const float m = __ldg(a.m_pre + e);
      x0 *= m;
      x1 *= m;
    }
    v[i] = make_float2(x0, x1);
  }
}

template <int M>
__device__ __forceinline__ void blue_reduce2(float& s0, float& s1, int t, int g, float* scr) {
  constexpr int T = BGeo<M>::T;
  constexpr int W = (T < 32) ? T : 32;
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if constexpr (T > 32) {
    float* s = scr + g * 32;
    if ((t & 31) == 0) {
      s[2 * (t >> 5)] = s0;
      s[2 * (t >> 5) + 1] = s1;
    }
    __syncthreads();
    if (t == 0) {
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int w = 0; w < T / 32; ++w) {
        sa += s[2 * w];
        sb += s[2 * w + 1];
      }
      s0 = sa;
      s1 = sb;
    }
  }
}

// y (stage-0 layout, complex = two traces) -> * m_post, store, intensity
template <int M>
__device__ __forceinline__ void blue_store(float2 (&v)[kE], const BlueArgs& a, int t, int g, bool act0, bool act1,
                                           int64_t p0, bool z0, bool z1, float* scr) {
  constexpr int T = BGeo<M>::T;
  float* r0 = a.out + p0 * a.n;
  float* r1 = r0 + a.n;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < kE; ++i) {
    const int e = t + i * T;
    if (e < a.n) {
      float y0 = z0 ? 0.f : v[i].x, y1 = z1 ? 0.f : v[i].y;
      if (a.m_post != nullptr) {
        const float m = __ldg(a.m_post + e);
        y0 *= m;
        y1 *= m;
      }
      if (act0) __stcs(r0 + e, y0);
      if (act1) __stcs(r1 + e, y1);
      s0 = fmaf(y0, y0, s0);
      s1 = fmaf(y1, y1, s1);
    }
  }
  if (a.img != nullptr) {
    blue_reduce2<M>(s0, s1, t, g, scr);
    if (t == 0) {
      if (act0) a.img[p0] = s0;
      if (act1) a.img[p0 + 1] = s1;
    }
  }
}

// fused chain for arbitrary n
template <int M, bool SPLIT = false>
__global__ void __launch_bounds__(BGeo<M>::NT, BGeo<M>::kMinBlocks) k_blue_fused(const BlueArgs a) {
  using GEO = BGeo<M>;
  constexpr int T = GEO::T, G = GEO::G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  float* scr = reinterpret_cast<float*>(smem + (size_t)G * padded_len(M));
  unsigned* nzbuf = reinterpret_cast<unsigned*>(scr + 32 * G);
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  float2* sm = smem + (size_t)g * padded_len(M);
  float2* stash = reinterpret_cast<float2*>(smem_raw + BGeo<M>::stash_off) + (size_t)g * M + t;   // SPLIT only
  const int64_t npairs = (a.P + 1) >> 1;
  const int64_t nitems = (npairs + G - 1) / G;
  int parity = 0;
  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x, parity ^= 1) {
    const int64_t p0 = (item * G + g) * 2;
    const bool act0 = p0 < a.P, act1 = p0 + 1 < a.P;
    float2 v[kE];
    bool nz0, nz1, z0, z1;
    blue_load<M>(v, a, t, act0, act1, p0, nz0, nz1);
    nz_publish<T>(nz0, nz1, t, g, parity, nzbuf, z0, z1);
    bluestein_dft<M, SPLIT>(v, t, sm, a, stash);            // X[k], k < n
#pragma unroll
    for (int i = 0; i < kE; ++i) {            // band-pass, then conj for the inverse DFT
      const int e = t + i * T;
      const float h = (e < a.n) ? __ldg(a.hn + e) : 0.f;
      v[i] = make_float2(v[i].x * h, -v[i].y * h);
    }
    bluestein_dft<M, SPLIT>(v, t, sm, a, stash);
#pragma unroll
    for (int i = 0; i < kE; ++i) v[i].y = -v[i].y;
    nz_resolve<T>(g, parity, nzbuf, z0, z1);
    blue_store<M>(v, a, t, g, act0, act1, p0, z0, z1, scr);
  }
}

// forward: spectra materialised for arbitrary n (math_tools::fft)
template <int M, bool SPLIT = false>
__global__ void __launch_bounds__(BGeo<M>::NT, BGeo<M>::kMinBlocks) k_blue_forward(const BlueArgs a) {
  using GEO = BGeo<M>;
  constexpr int T = GEO::T, G = GEO::G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  float* scr = reinterpret_cast<float*>(smem + (size_t)G * padded_len(M));
  unsigned* nzbuf = reinterpret_cast<unsigned*>(scr + 32 * G);
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  float2* sm = smem + (size_t)g * padded_len(M);
  float2* stash = reinterpret_cast<float2*>(smem_raw + BGeo<M>::stash_off) + (size_t)g * M + t;   // SPLIT only
  float* phs = reinterpret_cast<float*>(sm);
  const int64_t npairs = (a.P + 1) >> 1;
  const int64_t nitems = (npairs + G - 1) / G;
  const int n = a.n, F = n / 2 + 1;
  const int PF = F + 1;                        // phase buffer stride per trace (floats)
  const bool want_phase = a.phase != nullptr;
  int parity = 0;
  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x, parity ^= 1) {
    const int64_t p0 = (item * G + g) * 2;
    const bool act0 = p0 < a.P, act1 = p0 + 1 < a.P;
    float2 v[kE];
    bool nz0, nz1, z0, z1;
    blue_load<M>(v, a, t, act0, act1, p0, nz0, nz1);
    nz_publish<T>(nz0, nz1, t, g, parity, nzbuf, z0, z1);
    if (a.win != nullptr) {
#pragma unroll
      for (int i = 0; i < kE; ++i) {
        const int e = t + i * T;
        if (e < n) {
          if (act0) __stcs(a.win + p0 * n + e, v[i].x);
          if (act1) __stcs(a.win + (p0 + 1) * n + e, v[i].y);
        }
      }
    }
    bluestein_dft<M, SPLIT>(v, t, sm, a, stash);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kE; ++i) {
      const int e = t + i * T;
      if (e < n) sm[e] = v[i];                 // natural order, unpadded (strided only across i)
    }
    __syncthreads();
    nz_resolve<T>(g, parity, nzbuf, z0, z1);
    // split and write; raw phases go to registers first (the buffer is reused for them)
    float ph0[kE / 2 + 1], ph1[kE / 2 + 1];
#pragma unroll
    for (int u = 0; u < kE / 2 + 1; ++u) {
      const int k = t + u * T;
      ph0[u] = ph1[u] = 0.f;
      if (k < F) {
        const float2 za = sm[k];
        const float2 zb = sm[(n - k) % n];
        float2 x0 = make_float2(0.5f * (za.x + zb.x), 0.5f * (za.y - zb.y));
        float2 x1 = make_float2(0.5f * (za.y + zb.y), 0.5f * (zb.x - za.x));
        if (z0) x0 = make_float2(0.f, 0.f);
        if (z1) x1 = make_float2(0.f, 0.f);
        if (a.fft != nullptr) {
          if (act0) __stcs(a.fft + p0 * F + k, x0);
          if (act1) __stcs(a.fft + (p0 + 1) * F + k, x1);
        }
        if (a.amp != nullptr) {
          if (act0) __stcs(a.amp + p0 * F + k, sqrtf(fmaf(x0.x, x0.x, x0.y * x0.y)));
          if (act1) __stcs(a.amp + (p0 + 1) * F + k, sqrtf(fmaf(x1.x, x1.x, x1.y * x1.y)));
        }
        if (want_phase) {
          ph0[u] = atan2f(x0.y, x0.x);
          ph1[u] = atan2f(x1.y, x1.x);
        }
      }
    }
    if (want_phase) {
      __syncthreads();
#pragma unroll
      for (int u = 0; u < kE / 2 + 1; ++u) {
        const int k = t + u * T;
        if (k < F) {
          phs[k] = ph0[u];
          phs[PF + k] = ph1[u];
        }
      }
      __syncthreads();
      // threshold unwrap (src/math_tools.rs:224-237): the first two warps of the group scan one trace each,
      // every lane a contiguous chunk, warp-level exclusive scan of the chunk sums
      const int wid = t >> 5, lane = t & 31;
      if (wid < 2 || T < 64) {
        const int ntr = (T >= 64) ? 1 : 2;       // groups narrower than two warps: each scanning lane set does both
        for (int rr = 0; rr < ntr; ++rr) {
          const int r = (T >= 64) ? wid : rr;
          const int lanes = (T >= 32) ? 32 : T;
          const int ln = (T >= 32) ? lane : t;
          float* pr = phs + r * PF;
          const int chunk = (F - 1 + lanes - 1) / lanes;
          const int k0 = 1 + ln * chunk, k1 = min(F, k0 + chunk);
          const float kPi = 3.14159265358979323846f, kTwoPi = 2.0f * kPi;
          float run = 0.f;
          float prev = (k0 <= F) ? pr[min(k0, F) - 1] : 0.f;
          for (int k = k0; k < k1; ++k) {
            const float val = pr[k];
            float d = val - prev;
            if (d > kPi) d -= kTwoPi;
            else if (d < -kPi) d += kTwoPi;
            run += d;
            prev = val;
          }
          float inc = run;
          for (int o = 1; o < lanes; o <<= 1) {
            const float nb = __shfl_up_sync(0xffffffffu, inc, o, lanes);
            if (ln >= o) inc += nb;
          }
          float acc = pr[0] + (inc - run);
          __syncwarp();
          prev = (k0 <= F) ? pr[min(k0, F) - 1] : 0.f;
          // second sweep rewrites the chunk in place; the neighbour's `prev` was read above (warp-synchronous)
          float carry_prev = prev;
          __syncwarp();
          for (int k = k0; k < k1; ++k) {
            const float val = pr[k];
            float d = val - carry_prev;
            if (d > kPi) d -= kTwoPi;
            else if (d < -kPi) d += kTwoPi;
            acc += d;
            carry_prev = val;
            pr[k] = acc;
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < kE / 2 + 1; ++u) {
        const int k = t + u * T;
        if (k < F) {
          if (act0) __stcs(a.phase + p0 * F + k, phs[k]);
          if (act1) __stcs(a.phase + (p0 + 1) * F + k, phs[PF + k]);
        }
      }
    }
  }
}

// inverse: spectra in, arbitrary n (math_tools::ifft)
template <int M, bool SPLIT = false>
__global__ void __launch_bounds__(BGeo<M>::NT, BGeo<M>::kMinBlocks) k_blue_inverse(const BlueArgs a) {
  using GEO = BGeo<M>;
  constexpr int T = GEO::T, G = GEO::G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  float* scr = reinterpret_cast<float*>(smem + (size_t)G * padded_len(M));
  unsigned* nzbuf = reinterpret_cast<unsigned*>(scr + 32 * G);
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  float2* sm = smem + (size_t)g * padded_len(M);
  float2* stash = reinterpret_cast<float2*>(smem_raw + BGeo<M>::stash_off) + (size_t)g * M + t;   // SPLIT only
  const int64_t npairs = (a.P + 1) >> 1;
  const int64_t nitems = (npairs + G - 1) / G;
  const int n = a.n, F = n / 2 + 1;
  const float inv_n = 1.0f / (float)n;
  int parity = 0;
  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x, parity ^= 1) {
    const int64_t p0 = (item * G + g) * 2;
    const bool act0 = p0 < a.P, act1 = p0 + 1 < a.P;
    bool nz0 = false, nz1 = false, z0, z1;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < kE / 2 + 1; ++u) {
      const int k = t + u * T;
      if (k < F) {
        float2 x0 = act0 ? __ldcs(a.fft_in + p0 * F + k) : make_float2(0.f, 0.f);
        float2 x1 = act1 ? __ldcs(a.fft_in + (p0 + 1) * F + k) : make_float2(0.f, 0.f);
        const float s = (a.band != nullptr) ? __ldg(a.band + k) * inv_n : inv_n;
        x0.x *= s; x0.y *= s; x1.x *= s; x1.y *= s;
        nz0 |= (x0.x != 0.f) | (x0.y != 0.f);
        nz1 |= (x1.x != 0.f) | (x1.y != 0.f);
        const bool self_mirror = (k == 0) || (2 * k == n);   // c2r ignores the imaginary parts of DC / Nyquist
        // the inverse DFT is conj(DFT(conj Z)): store conj(Z) directly
        if (self_mirror) {
          sm[k] = make_float2(x0.x, -x1.x);
        } else {
          sm[k] = make_float2(x0.x - x1.y, -(x0.y + x1.x));
          sm[n - k] = make_float2(x0.x + x1.y, -(x1.x - x0.y));
        }
      }
    }
    nz_publish<T>(nz0, nz1, t, g, parity, nzbuf, z0, z1);
    __syncthreads();
    float2 v[kE];
#pragma unroll
    for (int i = 0; i < kE; ++i) {
      const int e = t + i * T;
      v[i] = (e < n) ? sm[e] : make_float2(0.f, 0.f);
    }
    bluestein_dft<M, SPLIT>(v, t, sm, a, stash);
#pragma unroll
    for (int i = 0; i < kE; ++i) v[i].y = -v[i].y;
    nz_resolve<T>(g, parity, nzbuf, z0, z1);
    blue_store<M>(v, a, t, g, act0, act1, p0, z0, z1, scr);
  }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
int blue_fft_size(int n) {
  int m = 64;
  while (m < 2 * n - 1) m <<= 1;
  return m;
}

bool blue_supported(int n) { return n >= 2 && 2 * n - 1 <= 16384; }
constexpr int kBlueSplitM = 16384;   // plans of this size run as two 8192-point sub-spectra

template <int M> static void blue_plan(int& ns, int (&r)[4]) {
  ns = Plan<M>::ns;
  for (int i = 0; i < 4; ++i) r[i] = Plan<M>::r[i];
}
static bool blue_plan_of(int m, int& ns, int (&r)[4]) {
  switch (m) {
    case 64: blue_plan<64>(ns, r); return true;
    case 128: blue_plan<128>(ns, r); return true;
    case 256: blue_plan<256>(ns, r); return true;
    case 512: blue_plan<512>(ns, r); return true;
    case 1024: blue_plan<1024>(ns, r); return true;
    case 2048: blue_plan<2048>(ns, r); return true;
    case 4096: blue_plan<4096>(ns, r); return true;
    case 8192: blue_plan<8192>(ns, r); return true;
    default: return false;
  }
}

// chirp[k] = exp(-i pi k^2 / n) (k^2 reduced mod 2n exactly), bhat = DFT_M(b) / M in register order with
// b[j] = conj(chirp[|j|]) for |j| < n placed circularly, hn[k] = band[min(k, n-k)] / n
int build_bluestein_tables(int n, const float* band, std::vector<float2>& chirp, std::vector<float2>& bhat,
                           std::vector<float>& hn, int& m_out) {
  const int m = blue_fft_size(n);
  int ns, r[4];
  const bool split = (m == kBlueSplitM);
  if (!blue_supported(n) || !blue_plan_of(split ? m / 2 : m, ns, r)) return THZ_EINVAL;
  m_out = m;
  std::vector<double> cr(n), ci(n);
  chirp.resize(n);
  for (int k = 0; k < n; ++k) {
    const long long q = ((long long)k * k) % (2LL * n);
    const double ang = -M_PI * (double)q / (double)n;
    cr[k] = cos(ang);
    ci[k] = sin(ang);
    chirp[k] = make_float2((float)cr[k], (float)ci[k]);
  }
  // b circular, then a plain O(M log M) radix-2 transform in double on the host (runs once per plan)
  std::vector<double> br(m, 0.0), bi(m, 0.0);
  for (int j = 0; j < n; ++j) {
    br[j] = cr[j];
    bi[j] = -ci[j];
    if (j) {
      br[m - j] = cr[j];
      bi[m - j] = -ci[j];
    }
  }
  {   // in-place iterative FFT (forward)
    for (int i = 1, j = 0; i < m; ++i) {
      int bit = m >> 1;
      for (; j & bit; bit >>= 1) j ^= bit;
      j ^= bit;
      if (i < j) {
        std::swap(br[i], br[j]);
        std::swap(bi[i], bi[j]);
      }
    }
    for (int len = 2; len <= m; len <<= 1) {
      const double ang = -2.0 * M_PI / len;
      for (int s = 0; s < m; s += len)
        for (int k = 0; k < len / 2; ++k) {
          const double wr = cos(ang * k), wi = sin(ang * k);
          const double ur = br[s + k], ui = bi[s + k];
          const double vr = br[s + k + len / 2] * wr - bi[s + k + len / 2] * wi;
          const double vi = br[s + k + len / 2] * wi + bi[s + k + len / 2] * wr;
          br[s + k] = ur + vr; bi[s + k] = ui + vi;
          br[s + k + len / 2] = ur - vr; bi[s + k + len / 2] = ui - vi;
        }
    }
  }
  const int RLs = r[ns - 1];
  if (!split) {
    bhat.assign(m, make_float2(0.f, 0.f));
    for (int beta = 0; beta < m / RLs; ++beta)
      for (int mm = 0; mm < RLs; ++mm) {
        int p = beta * RLs + mm, k = 0, w = 1, L = m;
        for (int s = 0; s < ns; ++s) {
          const int S = L / r[s];
          const int q = p / S;
          p -= q * S;
          k += q * w;
          w *= r[s];
          L = S;
        }
        bhat[(size_t)mm * (m / RLs) + beta] = make_float2((float)(br[k] / m), (float)(bi[k] / m));
      }
  } else {
    // even bins | odd bins, each in the register order of the (m/2)-point plan, then v[t] = exp(-2 pi i t / m), t < m/32
    const int mh = m / 2;
    bhat.assign((size_t)m + mh / kE, make_float2(0.f, 0.f));
    for (int beta = 0; beta < mh / RLs; ++beta)
      for (int mm = 0; mm < RLs; ++mm) {
        int p = beta * RLs + mm, j = 0, w = 1, L = mh;
        for (int s = 0; s < ns; ++s) {
          const int S = L / r[s];
          const int q = p / S;
          p -= q * S;
          j += q * w;
          w *= r[s];
          L = S;
        }
        const size_t o = (size_t)mm * (mh / RLs) + beta;
        bhat[o] = make_float2((float)(br[2 * j] / m), (float)(bi[2 * j] / m));
        bhat[(size_t)mh + o] = make_float2((float)(br[2 * j + 1] / m), (float)(bi[2 * j + 1] / m));
      }
    for (int t = 0; t < mh / kE; ++t) {
      const double ang = -2.0 * M_PI * (double)t / (double)m;
      bhat[(size_t)m + t] = make_float2((float)cos(ang), (float)sin(ang));
    }
  }
  hn.assign(n, 0.f);
  for (int k = 0; k < n; ++k) {
    const int kk = (k <= n / 2) ? k : n - k;
    hn[k] = (band ? band[kk] : 1.0f) / (float)n;
  }
  return THZ_OK;
}

template <int M, typename K>
static int launch_blue(thz_ctx* c, cudaStream_t s, K kernel, const BlueArgs& a, bool split = false) {
  using GEO = BGeo<M>;
  const size_t smem = split ? GEO::smem_bytes_split : GEO::smem_bytes;
  const void* key = (const void*)kernel;
  auto it = c->occ.find(key);
  if (it == c->occ.end()) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaFuncSetAttribute(bluestein)");
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, GEO::NT, smem);
    if (e != cudaSuccess || nb < 1) return cuda_fail(c, e, "occupancy(bluestein)");
    it = c->occ.emplace(key, nb).first;
  }
  const int64_t npairs = (a.P + 1) / 2;
  const int64_t nitems = (npairs + GEO::G - 1) / GEO::G;
  if (nitems <= 0) return THZ_OK;
  int64_t grid = (int64_t)c->sm_count * it->second;
  if (grid > nitems) grid = nitems;
  kernel<<<(unsigned)grid, GEO::NT, smem, s>>>(a);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "bluestein kernel launch");
  return THZ_OK;
}

template <int M> static int do_bfused(thz_ctx* c, cudaStream_t s, const BlueArgs& a) { return launch_blue<M>(c, s, k_blue_fused<M>, a); }
template <int M> static int do_bforward(thz_ctx* c, cudaStream_t s, const BlueArgs& a) { return launch_blue<M>(c, s, k_blue_forward<M>, a); }
template <int M> static int do_binverse(thz_ctx* c, cudaStream_t s, const BlueArgs& a) { return launch_blue<M>(c, s, k_blue_inverse<M>, a); }

template <int M> static int do_bfused_split(thz_ctx* c, cudaStream_t s, const BlueArgs& a) { return launch_blue<M>(c, s, k_blue_fused<M, true>, a, true); }
template <int M> static int do_bforward_split(thz_ctx* c, cudaStream_t s, const BlueArgs& a) { return launch_blue<M>(c, s, k_blue_forward<M, true>, a, true); }
template <int M> static int do_binverse_split(thz_ctx* c, cudaStream_t s, const BlueArgs& a) { return launch_blue<M>(c, s, k_blue_inverse<M, true>, a, true); }

#define THZ_DISPATCH_BM(m, FN, ...)                \
  switch (m) {                                     \
    case kBlueSplitM: return FN##_split<kBlueSplitM / 2>(__VA_ARGS__); \
    case 64: return FN<64>(__VA_ARGS__);           \
    case 128: return FN<128>(__VA_ARGS__);         \
    case 256: return FN<256>(__VA_ARGS__);         \
    case 512: return FN<512>(__VA_ARGS__);         \
    case 1024: return FN<1024>(__VA_ARGS__);       \
    case 2048: return FN<2048>(__VA_ARGS__);       \
    case 4096: return FN<4096>(__VA_ARGS__);       \
    case 8192: return FN<8192>(__VA_ARGS__);       \
    default: return THZ_EINVAL;                    \
  }

static int blue_base(thz_ctx* c, BlueArgs& a, int64_t P) {
  const TracePlan& p = c->plan;
  if (p.blue_m == 0) return set_err(c, THZ_ESTATE, "no Bluestein plan");
  const FftTables* tb = nullptr;
  int rc = get_tables(c, p.blue_m == kBlueSplitM ? kBlueSplitM / 2 : p.blue_m, &tb);
  if (rc != THZ_OK) return rc;
  a = BlueArgs{};
  a.tw = tb->d_tw;
  a.chirp = p.d_chirp;
  a.bhat = p.d_bhat;
  a.hn = p.d_hn;
  a.n = p.n;
  a.P = P;
  return THZ_OK;
}

int launch_blue_fused(thz_ctx* c, cudaStream_t s, const float* d_in, float* d_out, float* d_img, int64_t P) {
  BlueArgs a;
  int rc = blue_base(c, a, P);
  if (rc != THZ_OK) return rc;
  a.in = d_in; a.out = d_out; a.img = d_img;
  a.m_pre = c->plan.has_pre ? c->plan.d_m_pre : nullptr;
  a.m_post = c->plan.has_post ? c->plan.d_m_post : nullptr;
  THZ_DISPATCH_BM(c->plan.blue_m, do_bfused, c, s, a);
}

int launch_blue_forward(thz_ctx* c, cudaStream_t s, const float* d_in, float* d_win, float2* d_fft, float* d_amp,
                        float* d_phase, int64_t P) {
  BlueArgs a;
  int rc = blue_base(c, a, P);
  if (rc != THZ_OK) return rc;
  a.in = d_in; a.win = d_win; a.fft = d_fft; a.amp = d_amp; a.phase = d_phase;
  a.m_pre = c->plan.has_pre ? c->plan.d_m_pre : nullptr;
  THZ_DISPATCH_BM(c->plan.blue_m, do_bforward, c, s, a);
}

int launch_blue_inverse(thz_ctx* c, cudaStream_t s, const float2* d_fft, bool use_band, bool use_post, float* d_out,
                        float* d_img, int64_t P) {
  BlueArgs a;
  int rc = blue_base(c, a, P);
  if (rc != THZ_OK) return rc;
  a.fft_in = d_fft; a.out = d_out; a.img = d_img;
  a.band = (use_band && c->plan.has_band) ? c->plan.d_band : nullptr;
  a.m_post = (use_post && c->plan.has_post) ? c->plan.d_m_post : nullptr;
  THZ_DISPATCH_BM(c->plan.blue_m, do_binverse, c, s, a);
}

}  // namespace thz
