// thz_deconv.cu -- PSF-based frequency-dependent Richardson-Lucy deconvolution on sm_100a.
//
// Reference (paths under the upstream repository):
//   src/filters/deconvolution.rs:266-317, 574-609  convolve1d / filter_scan (FIR per trace)
//   src/filters/deconvolution.rs:620-712           richardson_lucy
//   src/filters/deconvolution.rs:432-545           direct_convolve2d / FFT convolve2d
//   src/filters/deconvolution.rs:963-1012, 1030    band energy, gains, band sum, intensity
//
// Restructuring (exact up to f32 rounding):
//  * The reference materialises one full band cube per FIR band.  Deconvolution is linear
//    per pixel, out[p] = sum_b g_b[p] (h_b * x[p]) = (sum_b g_b[p] h_b) * x[p], so the cube
//    is only streamed twice: pass A computes the B band-energy images, pass C applies the
//    per-pixel combined filter in the frequency domain.
//  * The FIR is symmetric (linear phase) and the reference keeps samples [249, 249+N) of the
//    full convolution, i.e. it applies the zero-phase centred filter: its spectrum is real.
//    A zero-padded transform of M >= N + 249 points makes the circular result identical.
//  * The PSF is an outer product (src/filters/psf.rs:305-311): each Richardson-Lucy 2-D
//    filtering is a row pass + a column pass on a TMA-staged tile (zero fill outside the
//    padded domain comes from the TMA out-of-bounds rule).
#include <cuda.h>
#include <cudaTypedefs.h>

// The FIR passes keep 130-190 KB of band tables hot in L1: fetch only the twiddle rows 1, 2, 4, 8 and form the
// other powers as products (measured: -2.5 % on the spectra pass, -5 % on the edge pass; the trace pass, whose
// tables fit anyway, is 2 % faster with the full table and keeps it)
#define THZ_TW_POWERS 1
#include "thz_deconv_dev.cuh"

#include <math.h>
#include <algorithm>
#include <type_traits>

namespace thz {

template <int M> struct DGeo {
  static constexpr int T = M / kE;
  static constexpr int NT = (T >= 256) ? T : 256;
  static constexpr int G = NT / T;
  static constexpr int kScr = (32 + kNzWords) * G;
  static constexpr size_t smem_bytes = (size_t)G * padded_len(M) * sizeof(float2) + kScr * sizeof(float);
  static constexpr int kMinBlocks = (NT == 256) ? 2 : 1;
};


template <int M>
__device__ __forceinline__ void dreduce2(float& a, float& b, int t, int g, float* scr) {
  constexpr int T = DGeo<M>::T;
  constexpr int W = (T < 32) ? T : 32;
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if constexpr (T > 32) {
    constexpr int NW = T / 32;
    float* s = scr + g * 32;   // reuse is separated by the barriers of the next transform
    if ((t & 31) == 0) {
      s[2 * (t >> 5)] = a;
      s[2 * (t >> 5) + 1] = b;
    }
    __syncthreads();
    if (t == 0) {
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) {
        sa += s[2 * w];
        sb += s[2 * w + 1];
      }
      a = sa;
      b = sb;
    }
  }
}

template <int M>
__device__ __forceinline__ void load_padded_pair(float2 (&v)[kE], const FirArgs& a, int t, bool act0, bool act1,
                                                 int64_t p0, bool& nz0, bool& nz1) {
  constexpr int T = DGeo<M>::T;
  const float* r0 = a.x + p0 * a.n;
  const float* r1 = r0 + a.n;
  nz0 = nz1 = false;
#pragma unroll
  for (int i = 0; i < kE; ++i) {
    const int e = t + i * T;
    const bool in = e < a.n;
    v[i].x = (act0 && in) ? __ldcs(r0 + e) : 0.f;
    v[i].y = (act1 && in) ? __ldcs(r1 + e) : 0.f;
    nz0 |= (v[i].x != 0.f);
    nz1 |= (v[i].y != 0.f);
  }
}

// ---- pass A: band energies ------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(DGeo<M>::NT, DGeo<M>::kMinBlocks) k_fir_energy(const FirArgs a) {
  using GEO = DGeo<M>;
  constexpr int T = GEO::T, G = GEO::G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  float* scr = reinterpret_cast<float*>(smem + (size_t)G * padded_len(M));
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  float2* sm = smem + (size_t)g * padded_len(M);
  const int64_t npairs = (a.P + 1) >> 1;
  const int64_t nitems = (npairs + G - 1) / G;
  constexpr int LAST = Plan<M>::ns - 1;
  constexpr int RL = Plan<M>::r[LAST];
  constexpr int UL = kE / RL;

  unsigned* nzbuf = reinterpret_cast<unsigned*>(scr + 32 * G);
  int parity = 0;
  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x, parity ^= 1) {
    const int64_t p0 = (item * G + g) * 2;
    const bool act0 = p0 < a.P, act1 = p0 + 1 < a.P;
    float2 z[kE];
    bool nz0, nz1, z0, z1;
    load_padded_pair<M>(z, a, t, act0, act1, p0, nz0, nz1);
    nz_publish<T>(nz0, nz1, t, g, parity, nzbuf, z0, z1);
    fft_forward<M>(z, t, sm, a.tw);
    nz_resolve<T>(g, parity, nzbuf, z0, z1);
    for (int b = 0; b < a.B; ++b) {
      const float* hq = a.hq + (size_t)b * M;
      float2 w[kE];
#pragma unroll
      for (int i = 0; i < kE; ++i) {
        const int u = i % UL, m = i / UL;
        const float h = __ldg(hq + m * (M / RL) + t + u * T);
        w[i] = make_float2(z[i].x * h, z[i].y * h);
      }
      fft_inverse<M>(w, t, sm, a.tw);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int i = 0; i < kE; ++i) {
        if (t + i * T < a.n) {
          s0 = fmaf(w[i].x, w[i].x, s0);
          s1 = fmaf(w[i].y, w[i].y, s1);
        }
      }
      dreduce2<M>(s0, s1, t, g, scr);
      if (t == 0) {
        // an all-zero trace has exactly zero band energy in the reference (-> NaN gain, quirk 10)
        if (act0) a.energy[(size_t)b * a.bstride + p0] = z0 ? 0.f : s0;
        if (act1) a.energy[(size_t)b * a.bstride + p0 + 1] = z1 ? 0.f : s1;
      }
    }
  }
}


// ---- pass A, fast form: Parseval total energy ------------------------------------------------
// E_b = sum_{k in [249, N+249)} y_b[k]^2 with y_b = h_b * x the FULL linear convolution (N + 498
// samples).  For M >= N + 498 the total energy is (1/M) sum_f |H_b[f]|^2 |X[f]|^2 (no circular
// aliasing), so one forward transform per pair and B weighted sums replace B inverse transforms;
// the 2 x 249 excluded samples are subtracted by k_fir_edges.  With two traces packed as
// Z = X1 + i X2 and W symmetric: for 0 < f < M/2, p = (|Z[f]|^2 + |Z[M-f]|^2)/2, c = Re(Z[f] Z[M-f]);
// for f in {0, M/2}, p = |Z|^2/2, c = (Re^2 - Im^2)/2; then E1 += W[f] (p + c), E2 += W[f] (p - c).
template <int M>
__global__ void __launch_bounds__(DGeo<M>::NT, DGeo<M>::kMinBlocks) k_fir_energy_total(const FirArgs a) {
  using GEO = DGeo<M>;
  constexpr int T = GEO::T, G = GEO::G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  float* scr = reinterpret_cast<float*>(smem + (size_t)G * padded_len(M));
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  float2* sm = smem + (size_t)g * padded_len(M);
  const int64_t npairs = (a.P + 1) >> 1;
  const int64_t nitems = (npairs + G - 1) / G;
  constexpr int LAST = Plan<M>::ns - 1;
  constexpr int RL = Plan<M>::r[LAST];
  constexpr int UL = kE / RL;
  constexpr int NLOW = kE / 2;            // registers with last-stage digit m < RL/2 hold bins < M/2
  constexpr int W = (T < 32) ? T : 32;
  constexpr int NW = (T + 31) / 32;
  static_assert(RL >= 2, "last stage radix");
  unsigned* nzbuf = reinterpret_cast<unsigned*>(scr + 32 * G);
  float* red = reinterpret_cast<float*>(sm);   // the transform buffer is free during the reduction
  int parity = 0;

  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x, parity ^= 1) {
    const int64_t p0 = (item * G + g) * 2;
    const bool act0 = p0 < a.P, act1 = p0 + 1 < a.P;
    float2 z[kE];
    bool nz0, nz1, z0, z1;
    load_padded_pair<M>(z, a, t, act0, act1, p0, nz0, nz1);
    nz_publish<T>(nz0, nz1, t, g, parity, nzbuf, z0, z1);
    fft_forward<M>(z, t, sm, a.tw);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kE; ++i) sm[pad_idx(pos_to_bin<M>(stage_elem<M, LAST>(t, i)))] = z[i];
    __syncthreads();
    nz_resolve<T>(g, parity, nzbuf, z0, z1);
    // per lower-half register: q1 = p + c, q2 = p - c
    float q1[NLOW], q2[NLOW];
#pragma unroll
    for (int j = 0; j < NLOW; ++j) {
      // lower-half registers are i = u + UL * m with m < RL/2, enumerated as j = u + UL * m
      const int i = j;                      // (u, m) -> u + UL*m keeps the same numbering for m < RL/2
      const int f = pos_to_bin<M>(stage_elem<M, LAST>(t, i));
      const float2 zz = z[i];
      if (f == 0) {
        q1[j] = zz.x * zz.x;
        q2[j] = zz.y * zz.y;
      } else {
        const float2 zp = sm[pad_idx(M - f)];
        const float p = 0.5f * (zz.x * zz.x + zz.y * zz.y + zp.x * zp.x + zp.y * zp.y);
        const float c = zz.x * zp.x - zz.y * zp.y;
        q1[j] = p + c;
        q2[j] = p - c;
      }
    }
    // Nyquist bin M/2 lives in register (u = 0, m = RL/2) of thread beta = 0
    float ny1 = 0.f, ny2 = 0.f;
    if (t == 0) {
      const float2 zz = z[UL * (RL / 2)];
      ny1 = zz.x * zz.x;
      ny2 = zz.y * zz.y;
    }
    __syncthreads();   // all partner reads done: the buffer becomes reduction scratch
    for (int b0 = 0; b0 < a.B; b0 += 4) {
      float e1[4] = {0.f, 0.f, 0.f, 0.f}, e2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const int b = b0 + bb;
        if (b < a.B) {
          const float* wq = a.wq + (size_t)b * (M / 2);
#pragma unroll
          for (int j = 0; j < NLOW; ++j) {
            const int u = j % UL, m = j / UL;
            const float w = __ldg(wq + m * (M / RL) + t + u * T);
            e1[bb] = fmaf(w, q1[j], e1[bb]);
            e2[bb] = fmaf(w, q2[j], e2[bb]);
          }
          if (t == 0) {
            const float w = __ldg(a.wnyq + b);
            e1[bb] = fmaf(w, ny1, e1[bb]);
            e2[bb] = fmaf(w, ny2, e2[bb]);
          }
        }
      }
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) {
          e1[bb] += __shfl_xor_sync(0xffffffffu, e1[bb], o);
          e2[bb] += __shfl_xor_sync(0xffffffffu, e2[bb], o);
        }
      }
      if constexpr (T > 32) {
        if ((t & 31) == 0) {
#pragma unroll
          for (int bb = 0; bb < 4; ++bb) {
            red[((t >> 5) * 4 + bb) * 2] = e1[bb];
            red[((t >> 5) * 4 + bb) * 2 + 1] = e2[bb];
          }
        }
        __syncthreads();
        if (t < 8) {   // thread t sums (band t/2, trace t%2) over the warps
          float acc = 0.f;
          for (int w = 0; w < NW; ++w) acc += red[(w * 4 + (t >> 1)) * 2 + (t & 1)];
          const int b = b0 + (t >> 1);
          const bool second = (t & 1) != 0;
          if (b < a.B && (second ? act1 : act0))
            a.energy[(size_t)b * a.bstride + p0 + (second ? 1 : 0)] = (second ? z1 : z0) ? 0.f : acc;
        }
        __syncthreads();
      } else {
        if (t == 0) {
#pragma unroll
          for (int bb = 0; bb < 4; ++bb) {
            const int b = b0 + bb;
            if (b < a.B) {
              if (act0) a.energy[(size_t)b * a.bstride + p0] = z0 ? 0.f : e1[bb];
              if (act1) a.energy[(size_t)b * a.bstride + p0 + 1] = z1 ? 0.f : e2[bb];
            }
          }
        }
      }
    }
  }
}


template <int N, bool STAGED>
__global__ void __launch_bounds__(SGeo<N>::NT, SGeo<N>::kMinBlocks) k_fir_energy_split(const FirArgs a) {
  using GEO = SGeo<N>;
  constexpr int T = GEO::T, G = GEO::G;
  constexpr int M = 2 * N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  float* scr = reinterpret_cast<float*>(smem + (size_t)G * padded_len(N));
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  float2* sm = smem + (size_t)g * padded_len(N);
  const int64_t npairs = (a.P + 1) >> 1;
  const int64_t nitems = (npairs + G - 1) / G;
  constexpr int LAST = Plan<N>::ns - 1;
  constexpr int RL = Plan<N>::r[LAST];
  constexpr int UL = kE / RL;
  constexpr int NLOW = kE / 2;
  unsigned* nzbuf = reinterpret_cast<unsigned*>(scr + 32 * G);
  float* red = reinterpret_cast<float*>(sm);
  int parity = 0;
  (void)M;
  SlabPipe<N> pipe;
  if constexpr (STAGED) pipe.init(smem_raw, a.x, a.P, nitems);

  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x, parity ^= 1) {
    const int64_t p0 = (item * G + g) * 2;
    const bool act0 = p0 < a.P, act1 = p0 + 1 < a.P;
    const float* slab = nullptr;
    if constexpr (STAGED) slab = pipe.acquire(item);
    float2 z[kE];
    bool nz0, nz1, z0, z1;
    float q1e[NLOW], q2e[NLOW], q1o[NLOW], q2o[NLOW];
    // even bins
    if constexpr (STAGED) load_pair_n<N>(z, slab, t, g, act0, act1, nz0, nz1);
    else load_pair_direct<N>(z, a.x, p0, t, act0, act1, nz0, nz1);   // (an L2 prefetch of the next pair costs 4 % here)
    nz_publish<T>(nz0, nz1, t, g, parity, nzbuf, z0, z1);
    fft_forward<N>(z, t, sm, a.tw);
    __syncthreads();
#pragma unroll
    for (int i = kE / 2; i < kE; ++i) sm[pad_idx(pos_to_bin<N>(stage_elem<N, LAST>(t, i)))] = z[i];   // mirrors of lower-half bins are upper-half registers
    __syncthreads();
    nz_resolve<T>(g, parity, nzbuf, z0, z1);
    // odd bins: the second read of the pair (L2) is issued before the even-bin terms are formed, so that its
    // latency overlaps them
    float2 zo[kE];
    bool d0, d1;
    if constexpr (STAGED) load_pair_n<N>(zo, slab, t, g, act0, act1, d0, d1);
    else load_pair_direct<N>(zo, a.x, p0, t, act0, act1, d0, d1);
    parseval_terms<N, false>(z, sm, t, q1e, q2e);
    float ny1 = 0.f, ny2 = 0.f;   // bin M/2 = even index N/2: register (u = 0, digit RL/2) of thread 0
    if (t == 0) {
      const float2 zz = z[UL * (RL / 2)];
      ny1 = zz.x * zz.x;
      ny2 = zz.y * zz.y;
    }
    modulate<N>(zo, a.mod, t);
    fft_forward<N>(zo, t, sm, a.tw);   // its first barrier orders the partner reads above before the exchange
    __syncthreads();
#pragma unroll
    for (int i = kE / 2; i < kE; ++i) sm[pad_idx(pos_to_bin<N>(stage_elem<N, LAST>(t, i)))] = zo[i];   // mirrors of lower-half bins are upper-half registers
    __syncthreads();
    parseval_terms<N, true>(zo, sm, t, q1o, q2o);
    __syncthreads();   // partner reads done: the buffer becomes reduction scratch
    band_energy_reduce<N>(a, q1e, q2e, q1o, q2o, ny1, ny2, t, red, p0, act0, act1, z0, z1);
  }
}

// Y = S Z + D conj(Z_mirror) for one sub-spectrum, S = sum_b (g0+g1)/2 H_b, D = sum_b (g0-g1)/2 H_b,
// in two batches of 8 registers to bound the register footprint
template <int N, bool ODD>
__device__ __forceinline__ void mix_subspectrum(float2 (&z)[kE], const float2* sm, int t, const FirArgs& a,
                                                const float* __restrict__ htab, int64_t p0, bool act0, bool act1,
                                                bool& bad0, bool& bad1, float gmul = 0.5f) {
  constexpr int T = SGeo<N>::T;
  constexpr int LAST = Plan<N>::ns - 1;
  constexpr int RL = Plan<N>::r[LAST];
  constexpr int UL = kE / RL;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float sacc[8], dacc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sacc[j] = dacc[j] = 0.f;
    for (int b = 0; b < a.B; ++b) {
      float g0 = act0 ? __ldg(a.gain + (size_t)b * a.bstride + p0) : 0.f;
      float g1 = act1 ? __ldg(a.gain + (size_t)b * a.bstride + p0 + 1) : 0.f;
      if (!(fabsf(g0) <= 3.0e38f)) { bad0 = true; g0 = 0.f; }
      if (!(fabsf(g1) <= 3.0e38f)) { bad1 = true; g1 = 0.f; }
      const float gs = gmul * (g0 + g1), gd = gmul * (g0 - g1);
      const float* hq = htab + (size_t)b * N;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = half * 8 + j;
        const int u = i % UL, m = i / UL;
        const float h = __ldg(hq + m * (N / RL) + t + u * T);
        sacc[j] = fmaf(gs, h, sacc[j]);
        dacc[j] = fmaf(gd, h, dacc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int i = half * 8 + j;
      const int k = pos_to_bin<N>(stage_elem<N, LAST>(t, i));
      const float2 zp = sm[pad_idx(ODD ? (N - 1 - k) : ((N - k) & (N - 1)))];
      z[i] = make_float2(fmaf(sacc[j], z[i].x, dacc[j] * zp.x), fmaf(sacc[j], z[i].y, -dacc[j] * zp.y));
    }
  }
}

template <int N>
__global__ void __launch_bounds__(SGeo<N>::NT, SGeo<N>::kMinBlocks) k_fir_apply_split(const FirArgs a) {
  using GEO = SGeo<N>;
  constexpr int T = GEO::T, G = GEO::G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  float* scr = reinterpret_cast<float*>(smem + (size_t)G * padded_len(N));
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  float2* sm = smem + (size_t)g * padded_len(N);
  const int64_t npairs = (a.P + 1) >> 1;
  const int64_t nitems = (npairs + G - 1) / G;
  constexpr int LAST = Plan<N>::ns - 1;
  SlabPipe<N> pipe;
  pipe.init(smem_raw, a.x, a.P, nitems);

  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int64_t p0 = (item * G + g) * 2;
    const bool act0 = p0 < a.P, act1 = p0 + 1 < a.P;
    const float* slab = pipe.acquire(item);
    bool nz0, nz1, bad0 = false, bad1 = false;
    float2 ye[kE];
    {   // even bins
      float2 z[kE];
      load_pair_n<N>(z, slab, t, g, act0, act1, nz0, nz1);
      fft_forward<N>(z, t, sm, a.tw);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kE; ++i) sm[pad_idx(pos_to_bin<N>(stage_elem<N, LAST>(t, i)))] = z[i];
      __syncthreads();
      mix_subspectrum<N, false>(z, sm, t, a, a.he, p0, act0, act1, bad0, bad1);
      fft_inverse<N>(z, t, sm, a.tw);
#pragma unroll
      for (int i = 0; i < kE; ++i) ye[i] = z[i];
    }
    float2 z[kE];
    load_pair_n<N>(z, slab, t, g, act0, act1, nz0, nz1);
    modulate<N>(z, a.mod, t);
    fft_forward<N>(z, t, sm, a.tw);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kE; ++i) sm[pad_idx(pos_to_bin<N>(stage_elem<N, LAST>(t, i)))] = z[i];
    __syncthreads();
    mix_subspectrum<N, true>(z, sm, t, a, a.ho, p0, act0, act1, bad0, bad1);
    fft_inverse<N>(z, t, sm, a.tw);
    // y[n] = ye[n] + conj(w_M^n) yo[n]   (1/M is folded into the H tables)
    const float kNaN = __int_as_float(0x7fc00000);
    float* r0 = a.out + p0 * N + t;
    float* r1 = r0 + N;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < kE; ++i) {
      const float2 w = __ldg(a.mod + t + i * T);
      const float2 yo = cmul_conj(z[i], w);
      const float y0 = bad0 ? kNaN : ye[i].x + yo.x, y1 = bad1 ? kNaN : ye[i].y + yo.y;
      if (act0) __stcs(r0 + i * T, y0);
      if (act1) __stcs(r1 + i * T, y1);
      s0 = fmaf(y0, y0, s0);
      s1 = fmaf(y1, y1, s1);
    }
    if (a.img != nullptr) {
      // group reduction (same scheme as the trace pass)
      constexpr int W = (T < 32) ? T : 32;
#pragma unroll
      for (int o = W / 2; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      }
      if constexpr (T > 32) {
        float* sc = scr + g * 32;
        if ((t & 31) == 0) {
          sc[2 * (t >> 5)] = s0;
          sc[2 * (t >> 5) + 1] = s1;
        }
        __syncthreads();
        if (t == 0) {
          float sa = 0.f, sb = 0.f;
#pragma unroll
          for (int w = 0; w < T / 32; ++w) {
            sa += sc[2 * w];
            sb += sc[2 * w + 1];
          }
          s0 = sa;
          s1 = sb;
        }
      }
      if (t == 0) {
        if (act0) a.img[p0] = s0;
        if (act1) a.img[p0 + 1] = s1;
      }
    }
  }
}

// ---- pass A, edges: subtract the energy of the 2 x 249 samples outside the "same" window -----------
// head: y[k], k < 249 = first 249 outputs of conv(x[0..249), h_b[0..249)); tail: last 249 outputs of
// conv(x[N-249..N), h_b[250..499)).  Both are length-497 linear convolutions: exact in a 512-point
// circular transform.  One group of 32 threads (one warp) per trace pair; both edges in sequence.
__global__ void __launch_bounds__(256, 2) k_fir_edges(const FirArgs a) {
  constexpr int M = 512;
  using GEO = DGeo<M>;
  constexpr int T = GEO::T, G = GEO::G;   // 32 threads per pair, 8 pairs per CTA
  static_assert(T == 32, "one warp per pair");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  float2* sm = smem + (size_t)g * padded_len(M);
  const int64_t npairs = (a.P + 1) >> 1;
  const int64_t nitems = (npairs + G - 1) / G;
  constexpr int LAST = Plan<M>::ns - 1;
  constexpr int RL = Plan<M>::r[LAST];
  constexpr int UL = kE / RL;
  constexpr int kSeg = (THZ_FIR_TAPS - 1) / 2;   // 249

  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int64_t p0 = (item * G + g) * 2;
    const bool act0 = p0 < a.P, act1 = p0 + 1 < a.P;
    // eight bands per round: acc[2*bb + trace] collects head + tail energy of band bg + bb over this lane's
    // outputs; one halving shuffle tree per round instead of a butterfly per band and edge
    for (int bg = 0; bg < a.B; bg += 8) {
      float acc[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] = 0.f;
      for (int edge = 0; edge < 2; ++edge) {
        const int off = edge ? a.n - kSeg : 0;
        const float* r0 = a.x + p0 * a.n + off;
        const float* r1 = r0 + a.n;
        float2 z[kE];
#pragma unroll
        for (int i = 0; i < kE; ++i) {
          const int e = t + i * T;
          const bool in = e < kSeg;
          z[i].x = (act0 && in) ? __ldg(r0 + e) : 0.f;
          z[i].y = (act1 && in) ? __ldg(r1 + e) : 0.f;
        }
        fft_forward<M>(z, t, sm, a.tw512);
        const int lo = edge ? kSeg - 1 : 0, hi = edge ? 2 * kSeg - 1 : kSeg;   // kept outputs [lo, hi)
#pragma unroll 1
        for (int bb = 0; bb < 8; ++bb) {   // rolled: acc is indexed at run time (a 64-byte local array, L1 resident)
          const int b = bg + bb;
          if (b < a.B) {
            const float2* hq = a.edge + ((size_t)edge * a.B + b) * M;
            float2 w[kE];
#pragma unroll
            for (int i = 0; i < kE; ++i) {
              const int u = i % UL, m = i / UL;
              w[i] = cmul(z[i], __ldg(hq + m * (M / RL) + t + u * T));
            }
            fft_inverse<M>(w, t, sm, a.tw512);
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int i = 0; i < kE; ++i) {
              const int e = t + i * T;
              if (e >= lo && e < hi) {
                s0 = fmaf(w[i].x, w[i].x, s0);
                s1 = fmaf(w[i].y, w[i].y, s1);
              }
            }
            acc[2 * bb] += s0;
            acc[2 * bb + 1] += s1;
          }
        }
      }
      // lane l ends with the warp total of value l >> 1 = (band bg + (l >> 2), trace (l >> 1) & 1)
#pragma unroll
      for (int half = 8, o = 16; half >= 1; half >>= 1, o >>= 1) {
        const bool up = (t & o) != 0;
#pragma unroll
        for (int k = 0; k < half; ++k) {
          const float send = up ? acc[k] : acc[k + half];
          const float keep = up ? acc[k + half] : acc[k];
          acc[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
      }
      acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], 1);
      const int v = t >> 1, b = bg + (v >> 1);
      const bool second = (v & 1) != 0;
      if ((t & 1) == 0 && b < a.B && (second ? act1 : act0)) {
        // only this warp touches the pair's entries; exact zeros (dead pixels) stay zero
        float* e = a.energy + (size_t)b * a.bstride + p0 + (second ? 1 : 0);
        const float cur = *e;
        if (cur != 0.f) *e = fmaxf(cur - acc[0], 0.f);
      }
    }
  }
}

// Paired form of the mix for a full N-point spectrum with a real, even H (zero-phase FIR): the bins k and
// N - k share S and D, so the owner of the lower-half register (k < N/2) forms S and D once and produces
// both outputs,  Y[k] = S Z[k] + D conj(Z[N-k])  (kept)  and  Y[N-k] = S Z[N-k] + D conj(Z[k])  (written over
// Z[N-k] in shared memory, which only this thread reads).  After a barrier every thread fetches its
// upper-half registers.  Halves the table loads and the S / D arithmetic.  `sm` holds the spectrum in
// natural order on entry; thread 0 also owns the self-mirrored Nyquist bin (register kE/2).
// lane-distributed gains of a pair: lane b (b < 16) holds gain[b][p0], lane 16 + b holds gain[b][p0 + 1]; one
// load per lane, issued before the forward transform so that its L2 latency is hidden, one live register.
// Non-finite gains (dead pixel: sqrt(u / 0), quirk 10) are reported through bad0 / bad1 and replaced by 0.
constexpr int kGainLanes = 16;
__device__ __forceinline__ float prefetch_pair_gains(const FirArgs& a, int64_t p0, bool act0, bool act1, bool& bad0,
                                                     bool& bad1) {
  const int lane = threadIdx.x & 31, b = lane & (kGainLanes - 1);
  const bool second = lane >= kGainLanes;
  float gv = 0.f;
  if (b < a.B && (second ? act1 : act0)) gv = __ldg(a.gain + (size_t)b * a.bstride + p0 + (second ? 1 : 0));
  const bool nonfinite = !(fabsf(gv) <= 3.0e38f);
  const unsigned m = __ballot_sync(0xffffffffu, nonfinite);
  bad0 = (m & 0xFFFFu) != 0u;
  bad1 = (m >> kGainLanes) != 0u;
  return nonfinite ? 0.f : gv;
}

template <int N>
__device__ __forceinline__ void mix_paired(float2 (&z)[kE], float2* sm, int t, const FirArgs& a,
                                           const float* __restrict__ htab, int64_t p0, bool act0, bool act1,
                                           bool& bad0, bool& bad1, float gmul, float gv) {
  constexpr int T = SGeo<N>::T;
  constexpr int LAST = Plan<N>::ns - 1;
  constexpr int RL = Plan<N>::r[LAST];
  constexpr int UL = kE / RL;
  constexpr int NLOW = kE / 2;
  float sacc[NLOW], dacc[NLOW];
#pragma unroll
  for (int j = 0; j < NLOW; ++j) sacc[j] = dacc[j] = 0.f;
  float sny = 0.f, dny = 0.f;
  (void)htab;
  for (int b0 = 0; b0 < a.B; b0 += 4) {
    float gs[4], gd[4];
#pragma unroll
    for (int bb = 0; bb < 4; ++bb) {
      const int b = b0 + bb;
      float g0, g1;
      if (b0 + 3 < kGainLanes) {   // uniform: this group of four comes from the prefetched lanes
        g0 = __shfl_sync(0xffffffffu, gv, b);
        g1 = __shfl_sync(0xffffffffu, gv, kGainLanes + b);
      } else {
        g0 = (act0 && b < a.B) ? __ldg(a.gain + (size_t)b * a.bstride + p0) : 0.f;
        g1 = (act1 && b < a.B) ? __ldg(a.gain + (size_t)b * a.bstride + p0 + 1) : 0.f;
        if (!(fabsf(g0) <= 3.0e38f)) { bad0 = true; g0 = 0.f; }
        if (!(fabsf(g1) <= 3.0e38f)) { bad1 = true; g1 = 0.f; }
      }
      gs[bb] = gmul * (g0 + g1);
      gd[bb] = gmul * (g0 - g1);
    }
#pragma unroll
    for (int j = 0; j < NLOW; ++j) {
      const int u = j % UL, m = j / UL;
      const float4 h4 = __ldg(reinterpret_cast<const float4*>(
          a.he4 + ((size_t)(b0 >> 2) * (N / 2) + (m * (N / RL) + t + u * T)) * 4));
      const float h[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        sacc[j] = fmaf(gs[bb], h[bb], sacc[j]);
        dacc[j] = fmaf(gd[bb], h[bb], dacc[j]);
      }
    }
    if (t == 0) {
      const float4 h4 = __ldg(reinterpret_cast<const float4*>(a.hny4 + b0));
      const float h[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        sny = fmaf(gs[bb], h[bb], sny);
        dny = fmaf(gd[bb], h[bb], dny);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NLOW; ++j) {
    const int k = pos_to_bin<N>(stage_elem<N, LAST>(t, j));
    const float2 zk = z[j];
    if (k == 0) {
      z[j] = make_float2((sacc[j] + dacc[j]) * zk.x, (sacc[j] - dacc[j]) * zk.y);
    } else {
      float2* mp = sm + pad_idx(N - k);
      const float2 zm = *mp;
      z[j] = make_float2(fmaf(sacc[j], zk.x, dacc[j] * zm.x), fmaf(sacc[j], zk.y, -dacc[j] * zm.y));
      *mp = make_float2(fmaf(sacc[j], zm.x, dacc[j] * zk.x), fmaf(sacc[j], zm.y, -dacc[j] * zk.y));
    }
  }
  if (t == 0) {
    float2* np = sm + pad_idx(N / 2);
    const float2 zn = *np;
    *np = make_float2((sny + dny) * zn.x, (sny - dny) * zn.y);
  }
  __syncthreads();
#pragma unroll
  for (int i = NLOW; i < kE; ++i) z[i] = sm[pad_idx(pos_to_bin<N>(stage_elem<N, LAST>(t, i)))];
}

// ---- pass C, circular form -------------------------------------------------------------------------
// With the zero-phase FIR (support |k| <= 249) and N >= 512 the N-point CIRCULAR convolution differs
// from the reference's linear "same" convolution only in the first and the last 249 outputs, where the
// pieces that the linear convolution pushes outside the window wrap around:
//     circ[n]           = same[n]           + tail[n]   (n < 249)
//     circ[N - 249 + m] = same[N - 249 + m] + head[m]   (m < 249)
// head / tail are the same two length-497 convolutions the energy pass subtracts (k_fir_edges); they are
// linear in the taps, so for the per-pixel filter sum_b g_b h_b one 512-point transform pair per edge gives
// them exactly.  Pass C is then ONE forward and ONE inverse N-point transform per trace pair
// (k_fir_apply_circ) plus the two small edge transforms (k_fir_edge_corr) instead of the four N-point
// transforms of the zero-padded split form.
//
// k_fir_edge_corr: one warp per trace pair and edge; writes corr[pair][edge][256] (float2: the two traces
// of the pair), edge 0 = head (applies to outputs N-249+m), edge 1 = tail (applies to outputs n).
constexpr int kCorrStride = 256;
constexpr int64_t kCorrChunkPairs = 131072;   // 512 MiB of corrections per chunk

__global__ void __launch_bounds__(256, 2) k_fir_edge_corr(const FirArgs a) {
  constexpr int M = 512;
  using GEO = DGeo<M>;
  constexpr int T = GEO::T, G = GEO::G;
  static_assert(T == 32, "one warp per pair");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  float2* sm = smem + (size_t)g * padded_len(M);
  const int64_t npairs = (a.P + 1) >> 1;
  const int64_t nitems = (npairs + G - 1) / G;
  constexpr int LAST = Plan<M>::ns - 1;
  constexpr int RL = Plan<M>::r[LAST];
  constexpr int UL = kE / RL;
  constexpr int kSeg = (THZ_FIR_TAPS - 1) / 2;   // 249

  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int64_t pair = item * G + g;
    if (pair >= npairs) continue;                 // warps are independent (warp-scope barriers only)
    const int64_t p0 = pair * 2;
    const bool act1 = p0 + 1 < a.P;
    // pull the NEXT pair's edge samples into L2 (2 traces x 2 edges x 8 lines of 128 bytes = one line per lane): the
    // 16 warps of an SM cannot hide the DRAM latency of their own first loads
    {
      const int64_t npair = pair + (int64_t)gridDim.x * G;
      if (npair < npairs) {
        const int tr = t >> 4, ed = (t >> 3) & 1, ln = t & 7;
        const int64_t np = npair * 2 + tr;
        if (np < a.P) {
          const float* q = a.x + np * a.n + (ed ? a.n - 256 : 0) + ln * 32;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
        }
      }
    }
    // the pair's gains of the first sixteen bands, one per lane (lane b: trace 0, lane 16 + b: trace 1), fetched
    // before the transforms and broadcast by shuffles where the mix needs them: their L2 latency was the kernel's
    // largest stall (long scoreboard 3.5 per issue slot in the round-2 ncu capture); one live register
    constexpr int kPre = 16;
    float gv = 0.f;
    {
      const int b = t & (kPre - 1);
      const bool second = t >= kPre;
      if (b < a.B && (!second || act1)) gv = __ldg(a.gain + (size_t)b * a.bstride + p0 + (second ? 1 : 0));
      if (!(fabsf(gv) <= 3.0e38f)) gv = 0.f;   // the main kernel marks such traces NaN
    }
    for (int edge = 0; edge < 2; ++edge) {
      const int off = edge ? a.n - kSeg : 0;
      const float* r0 = a.x + p0 * a.n + off;
      const float* r1 = r0 + a.n;
      float2 z[kE];
#pragma unroll
      for (int i = 0; i < kE; ++i) {
        const int e = t + i * T;
        const bool in = e < kSeg;
        z[i].x = in ? __ldg(r0 + e) : 0.f;
        z[i].y = (act1 && in) ? __ldg(r1 + e) : 0.f;
      }
      fft_forward<M>(z, t, sm, a.tw512);
      __syncwarp();
#pragma unroll
      for (int i = kE / 2; i < kE; ++i) sm[pad_idx(pos_to_bin<M>(stage_elem<M, LAST>(t, i)))] = z[i];
      __syncwarp();
      // Y = S Z + D conj(Z_mirror), S = sum_b (g0+g1)/2 E_b, D = sum_b (g0-g1)/2 E_b (complex edge spectra).
      // The taps are real, E_b[M-k] = conj(E_b[k]): the owner of the lower-half register forms S and D once
      // and produces both Y[k] and Y[M-k] (see mix_paired); thread 0 also owns the Nyquist bin.
      {
        constexpr int NLOW = kE / 2;
        float2 sacc[NLOW], dacc[NLOW];
#pragma unroll
        for (int j = 0; j < NLOW; ++j) sacc[j] = dacc[j] = make_float2(0.f, 0.f);
        float2 sny = make_float2(0.f, 0.f), dny = make_float2(0.f, 0.f);
        auto add_band = [&](int b, float gs, float gd) {
          const float2* hq = a.edge + ((size_t)edge * a.B + b) * M;
#pragma unroll
          for (int j = 0; j < NLOW; ++j) {
            const int u = j % UL, m = j / UL;
            const float2 h = __ldg(hq + m * (M / RL) + t + u * T);
            sacc[j].x = fmaf(gs, h.x, sacc[j].x);
            sacc[j].y = fmaf(gs, h.y, sacc[j].y);
            dacc[j].x = fmaf(gd, h.x, dacc[j].x);
            dacc[j].y = fmaf(gd, h.y, dacc[j].y);
          }
          if (t == 0) {
            const float2 h = __ldg(hq + (RL / 2) * (M / RL));
            sny.x = fmaf(gs, h.x, sny.x);
            sny.y = fmaf(gs, h.y, sny.y);
            dny.x = fmaf(gd, h.x, dny.x);
            dny.y = fmaf(gd, h.y, dny.y);
          }
        };
        for (int b = 0; b < kPre && b < a.B; ++b) {
          const float g0 = __shfl_sync(0xffffffffu, gv, b), g1 = __shfl_sync(0xffffffffu, gv, kPre + b);
          add_band(b, 0.5f * (g0 + g1), 0.5f * (g0 - g1));
        }
        for (int b = kPre; b < a.B; ++b) {
          float g0 = __ldg(a.gain + (size_t)b * a.bstride + p0);
          float g1 = act1 ? __ldg(a.gain + (size_t)b * a.bstride + p0 + 1) : 0.f;
          if (!(fabsf(g0) <= 3.0e38f)) g0 = 0.f;
          if (!(fabsf(g1) <= 3.0e38f)) g1 = 0.f;
          add_band(b, 0.5f * (g0 + g1), 0.5f * (g0 - g1));
        }
        // S a + D conj(b)
        auto mixc = [](float2 S, float2 D, float2 za, float2 zb) {
          return make_float2(S.x * za.x - S.y * za.y + D.x * zb.x + D.y * zb.y,
                             S.x * za.y + S.y * za.x - D.x * zb.y + D.y * zb.x);
        };
#pragma unroll
        for (int j = 0; j < NLOW; ++j) {
          const int k = pos_to_bin<M>(stage_elem<M, LAST>(t, j));
          const float2 zk = z[j];
          if (k == 0) {
            z[j] = mixc(sacc[j], dacc[j], zk, zk);
          } else {
            float2* mp = sm + pad_idx(M - k);
            const float2 zm = *mp;
            z[j] = mixc(sacc[j], dacc[j], zk, zm);
            *mp = mixc(make_float2(sacc[j].x, -sacc[j].y), make_float2(dacc[j].x, -dacc[j].y), zm, zk);
          }
        }
        if (t == 0) {
          float2* np = sm + pad_idx(M / 2);
          const float2 zn = *np;
          *np = mixc(sny, dny, zn, zn);
        }
        __syncwarp();
#pragma unroll
        for (int i = NLOW; i < kE; ++i) z[i] = sm[pad_idx(pos_to_bin<M>(stage_elem<M, LAST>(t, i)))];
      }
      fft_inverse<M>(z, t, sm, a.tw512);
      const int lo = edge ? kSeg - 1 : 0;          // kept outputs [lo, lo + 249)
      float2* dst = a.corr + ((size_t)pair * 2 + edge) * kCorrStride;
#pragma unroll
      for (int i = 0; i < kE; ++i) {
        const int e = t + i * T - lo;
        if (e >= 0 && e < kSeg) dst[e] = z[i];
      }
      __syncwarp();   // the next edge's first exchange follows the reads of the inverse transform
    }
  }
}

template <int N, bool STAGED, bool SPEC = false>
__global__ void __launch_bounds__(SGeo<N>::NT, SGeo<N>::kMinBlocks) k_fir_apply_circ(const FirArgs a) {
  using GEO = SGeo<N>;
  constexpr int T = GEO::T, G = GEO::G;
  constexpr int kSeg = (THZ_FIR_TAPS - 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  float* scr = reinterpret_cast<float*>(smem + (size_t)G * padded_len(N));
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  float2* sm = smem + (size_t)g * padded_len(N);
  const int64_t npairs = (a.P + 1) >> 1;
  const int64_t nitems = (npairs + G - 1) / G;
  constexpr int LAST = Plan<N>::ns - 1;
  SlabPipe<N> pipe;
  if constexpr (STAGED) pipe.init(smem_raw, a.x, a.P, nitems);

  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int64_t pair = item * G + g;
    const int64_t p0 = pair * 2;
    const bool act0 = p0 < a.P, act1 = p0 + 1 < a.P;
    bool nz0, nz1, bad0 = false, bad1 = false;
    float2 z[kE];
    if constexpr (SPEC) {
      // spectral hand-off: the fused trace + energy kernel left FFT_N of the pair here, in register order
      const int64_t next = item + gridDim.x;
      if (next < nitems) {
        int64_t cnt = a.P - next * G * 2;
        if (cnt > 2 * G) cnt = 2 * G;
        prefetch_l2_slab(reinterpret_cast<const float*>(a.xspec) + next * G * 2 * N, cnt * N);
      }
      const float2* sp = a.xspec + pair * N + t;
      nz0 = nz1 = false;
#pragma unroll
      for (int i = 0; i < kE; ++i) z[i] = act0 ? __ldcs(sp + i * T) : make_float2(0.f, 0.f);
    } else if constexpr (STAGED) {
      const float* slab = pipe.acquire(item);
      load_pair_n<N>(z, slab, t, g, act0, act1, nz0, nz1);
    } else {
      const int64_t next = item + gridDim.x;
      if (next < nitems) {
        int64_t cnt = a.P - next * G * 2;
        if (cnt > 2 * G) cnt = 2 * G;
        prefetch_l2_slab(a.x + next * G * 2 * N, cnt * N);
      }
      load_pair_direct<N>(z, a.x, p0, t, act0, act1, nz0, nz1);
    }
    const float gv = prefetch_pair_gains(a, p0, act0, act1, bad0, bad1);
    if constexpr (!SPEC) fft_forward<N>(z, t, sm, a.tw);
    __syncthreads();   // SPEC: the previous item's exchanges are done with the buffer
#pragma unroll
    for (int i = kE / 2; i < kE; ++i) sm[pad_idx(pos_to_bin<N>(stage_elem<N, LAST>(t, i)))] = z[i];   // mirrors of lower-half bins are upper-half registers
    __syncthreads();
    // he holds H / (2N): the N-point inverse needs H / N
    mix_paired<N>(z, sm, t, a, a.he, p0, act0, act1, bad0, bad1, 1.0f, gv);
    // wrap-around corrections of this thread's outputs, in flight across the inverse transform
    const float2* cp = a.corr + (size_t)pair * 2 * kCorrStride;
    float2 cr[kE];
#pragma unroll
    for (int i = 0; i < kE; ++i) {
      cr[i] = make_float2(0.f, 0.f);
      if (T >= kSeg && i != 0 && i != kE - 1) continue;   // only the first / last register can be an edge
      const int n = t + i * T;
      if (act0) {
        if (n < kSeg) cr[i] = __ldg(cp + kCorrStride + n);
        else if (n >= N - kSeg) cr[i] = __ldg(cp + (n - (N - kSeg)));
      }
    }
    fft_inverse<N>(z, t, sm, a.tw);
    const float kNaN = __int_as_float(0x7fc00000);
    float* r0 = a.out + p0 * N + t;
    float* r1 = r0 + N;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < kE; ++i) {
      const float y0 = bad0 ? kNaN : z[i].x - cr[i].x, y1 = bad1 ? kNaN : z[i].y - cr[i].y;
      if (act0) __stcs(r0 + i * T, y0);
      if (act1) __stcs(r1 + i * T, y1);
      s0 = fmaf(y0, y0, s0);
      s1 = fmaf(y1, y1, s1);
    }
    if (a.img != nullptr) {
      constexpr int W = (T < 32) ? T : 32;
#pragma unroll
      for (int o = W / 2; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      }
      if constexpr (T > 32) {
        float* sc = scr + g * 32;
        if ((t & 31) == 0) {
          sc[2 * (t >> 5)] = s0;
          sc[2 * (t >> 5) + 1] = s1;
        }
        __syncthreads();
        if (t == 0) {
          float sa = 0.f, sb = 0.f;
#pragma unroll
          for (int w = 0; w < T / 32; ++w) {
            sa += sc[2 * w];
            sb += sc[2 * w + 1];
          }
          s0 = sa;
          s1 = sb;
        }
      }
      if (t == 0) {
        if (act0) a.img[p0] = s0;
        if (act1) a.img[p0 + 1] = s1;
      }
    }
  }
}

// ---- pass C: per-pixel combined filter sum_b g_b[p] h_b ------------------------------------
template <int M>
__global__ void __launch_bounds__(DGeo<M>::NT, DGeo<M>::kMinBlocks) k_fir_apply(const FirArgs a) {
  using GEO = DGeo<M>;
  constexpr int T = GEO::T, G = GEO::G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  float* scr = reinterpret_cast<float*>(smem + (size_t)G * padded_len(M));
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  float2* sm = smem + (size_t)g * padded_len(M);
  const int64_t npairs = (a.P + 1) >> 1;
  const int64_t nitems = (npairs + G - 1) / G;
  constexpr int LAST = Plan<M>::ns - 1;
  constexpr int RL = Plan<M>::r[LAST];
  constexpr int UL = kE / RL;

  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int64_t p0 = (item * G + g) * 2;
    const bool act0 = p0 < a.P, act1 = p0 + 1 < a.P;
    float2 z[kE];
    bool nz0, nz1;
    load_padded_pair<M>(z, a, t, act0, act1, p0, nz0, nz1);
    fft_forward<M>(z, t, sm, a.tw);
    // natural-order copy so that every thread can fetch the mirror bin Z[M - k]
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kE; ++i) sm[pad_idx(pos_to_bin<M>(stage_elem<M, LAST>(t, i)))] = z[i];
    __syncthreads();
    // S = (G0 + G1) / 2, D = (G0 - G1) / 2 with G_r[k] = sum_b gain[b][p_r] H_b[k] / M
    float sacc[kE], dacc[kE];
#pragma unroll
    for (int i = 0; i < kE; ++i) sacc[i] = dacc[i] = 0.f;
    bool bad0 = false, bad1 = false;   // non-finite gain: the reference's output trace is NaN (quirk 10)
    for (int b = 0; b < a.B; ++b) {
      float g0 = act0 ? __ldg(a.gain + (size_t)b * a.bstride + p0) : 0.f;
      float g1 = act1 ? __ldg(a.gain + (size_t)b * a.bstride + p0 + 1) : 0.f;
      if (!(fabsf(g0) <= 3.0e38f)) { bad0 = true; g0 = 0.f; }
      if (!(fabsf(g1) <= 3.0e38f)) { bad1 = true; g1 = 0.f; }
      const float gs = 0.5f * (g0 + g1), gd = 0.5f * (g0 - g1);
      const float* hq = a.hq + (size_t)b * M;
#pragma unroll
      for (int i = 0; i < kE; ++i) {
        const int u = i % UL, m = i / UL;
        const float h = __ldg(hq + m * (M / RL) + t + u * T);
        sacc[i] = fmaf(gs, h, sacc[i]);
        dacc[i] = fmaf(gd, h, dacc[i]);
      }
    }
    // Y[k] = S Z[k] + D conj(Z[M-k])
#pragma unroll
    for (int i = 0; i < kE; ++i) {
      const int k = pos_to_bin<M>(stage_elem<M, LAST>(t, i));
      const float2 zp = sm[pad_idx((M - k) & (M - 1))];
      z[i] = make_float2(fmaf(sacc[i], z[i].x, dacc[i] * zp.x), fmaf(sacc[i], z[i].y, -dacc[i] * zp.y));
    }
    fft_inverse<M>(z, t, sm, a.tw);
    const float kNaN = __int_as_float(0x7fc00000);
    float* r0 = a.out + p0 * a.n;
    float* r1 = r0 + a.n;
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < kE; ++i) {
      const int e = t + i * T;
      if (e < a.n) {
        const float y0 = bad0 ? kNaN : z[i].x, y1 = bad1 ? kNaN : z[i].y;
        if (act0) __stcs(r0 + e, y0);
        if (act1) __stcs(r1 + e, y1);
        s0 = fmaf(y0, y0, s0);
        s1 = fmaf(y1, y1, s1);
      }
    }
    if (a.img != nullptr) {
      dreduce2<M>(s0, s1, t, g, scr);
      if (t == 0) {
        if (act0) a.img[p0] = s0;
        if (act1) a.img[p0 + 1] = s1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
static int fir_fft_size(int n) {
  int m = 64;
  while (m < n + (THZ_FIR_TAPS - 1) / 2) m <<= 1;   // circular result exact for M >= N + 249
  return m;
}

template <int M> static void host_plan_m(int& ns, int (&r)[4]) {
  ns = Plan<M>::ns;
  for (int i = 0; i < 4; ++i) r[i] = Plan<M>::r[i];
}
static bool plan_of_m(int m, int& ns, int (&r)[4]) {
  switch (m) {
    case 64: host_plan_m<64>(ns, r); return true;
    case 128: host_plan_m<128>(ns, r); return true;
    case 256: host_plan_m<256>(ns, r); return true;
    case 512: host_plan_m<512>(ns, r); return true;
    case 1024: host_plan_m<1024>(ns, r); return true;
    case 2048: host_plan_m<2048>(ns, r); return true;
    case 4096: host_plan_m<4096>(ns, r); return true;
    case 8192: host_plan_m<8192>(ns, r); return true;
    default: return false;
  }
}

// zero-phase spectrum of the centred FIR at M points, / M, in last-stage register order
static int build_fir_hq(int m, const float* fir, std::vector<float>& hq) {
  int ns, r[4];
  if (!plan_of_m(m, ns, r)) return THZ_EINVAL;
  const int taps = THZ_FIR_TAPS, half = (taps - 1) / 2;
  std::vector<double> H(m / 2 + 1), ctab(m);
  for (int i = 0; i < m; ++i) ctab[i] = cos(2.0 * M_PI * (double)i / (double)m);
  for (int k = 0; k <= m / 2; ++k) {
    double acc = (double)fir[half];
    for (int j = 1; j <= half; ++j) {
      // taps are symmetric up to f32 rounding; use both sides so that the real part is exact
      acc += ((double)fir[half + j] + (double)fir[half - j]) * ctab[(int)(((long)j * k) & (m - 1))];
    }
    H[k] = acc;
  }
  const int RLs = r[ns - 1];
  hq.assign(m, 0.f);
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
      const int kk = (k <= m / 2) ? k : m - k;
      hq[(size_t)mm * (m / RLs) + beta] = (float)(H[kk] / (double)m);
    }
  return THZ_OK;
}

// the same spectrum in natural bin order (k in [0, m)), for transform sizes without a monolithic plan
static void fir_spectrum_natural(int m, const float* fir, std::vector<float>& hnat) {
  const int taps = THZ_FIR_TAPS, half = (taps - 1) / 2;
  std::vector<double> ctab(m);
  for (int i = 0; i < m; ++i) ctab[i] = cos(2.0 * M_PI * (double)i / (double)m);
  hnat.assign(m, 0.f);
  for (int k = 0; k <= m / 2; ++k) {
    double acc = (double)fir[half];
    for (int j = 1; j <= half; ++j)
      acc += ((double)fir[half + j] + (double)fir[half - j]) * ctab[(int)(((long)j * k) & (m - 1))];
    const float v = (float)(acc / (double)m);
    hnat[k] = v;
    if (k != 0 && k != m / 2) hnat[m - k] = v;
  }
}

struct FirTables {
  bool mono = true;        // a monolithic m-point plan exists (hq / wq tables are filled)
  int m = 0, B = 0;
  float* d_hq = nullptr;   // workspace slot WS_FIR, cached across calls while the taps do not change
  float* d_wq = nullptr;   // [B][m/2]
  float* d_wnyq = nullptr; // [B]
  float2* d_edge = nullptr;// [2][B][512]
  // split form (n_half = m/2 = trace length): he, ho [B][n_half]; we, wo [B][n_half/2]; mod [n_half] float2
  bool split = false;
  float *d_he = nullptr, *d_ho = nullptr, *d_we = nullptr, *d_wo = nullptr;
  float2* d_mod = nullptr;
  int Bp = 0;
  float *d_we4 = nullptr, *d_wo4 = nullptr, *d_he4 = nullptr, *d_hny4 = nullptr;   // band-interleaved
};

// spectra / 512 of the first (edge 0) and last (edge 1) 249 taps at 512 points, register order of Plan<512>
static void build_edge_tables(const float* fir, std::vector<float2>& head, std::vector<float2>& tail) {
  const int M = 512, seg = (THZ_FIR_TAPS - 1) / 2;
  int ns, r[4];
  plan_of_m(M, ns, r);
  const int RLs = r[ns - 1];
  std::vector<double> ct(M), st(M);
  for (int i = 0; i < M; ++i) {
    ct[i] = cos(2.0 * M_PI * i / M);
    st[i] = sin(2.0 * M_PI * i / M);
  }
  head.assign(M, make_float2(0.f, 0.f));
  tail.assign(M, make_float2(0.f, 0.f));
  for (int beta = 0; beta < M / RLs; ++beta)
    for (int mm = 0; mm < RLs; ++mm) {
      int p = beta * RLs + mm, k = 0, w = 1, L = M;
      for (int s2 = 0; s2 < ns; ++s2) {
        const int S = L / r[s2];
        const int q = p / S;
        p -= q * S;
        k += q * w;
        w *= r[s2];
        L = S;
      }
      double hr = 0, hi = 0, tr = 0, ti = 0;
      for (int j = 0; j < seg; ++j) {
        const int idx = (int)(((long)j * k) & (M - 1));
        hr += (double)fir[j] * ct[idx];
        hi -= (double)fir[j] * st[idx];
        tr += (double)fir[seg + 1 + j] * ct[idx];
        ti -= (double)fir[seg + 1 + j] * st[idx];
      }
      const size_t o = (size_t)mm * (M / RLs) + beta;
      head[o] = make_float2((float)(hr / M), (float)(hi / M));
      tail[o] = make_float2((float)(tr / M), (float)(ti / M));
    }
}

static uint64_t fnv1a(const void* p, size_t n, uint64_t h) {
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < n; ++i) {
    h ^= b[i];
    h *= 0x100000001B3ull;
  }
  return h;
}

static int upload_fir_tables(thz_ctx* c, cudaStream_t s, int n, const thz_band_plan* bands, int B, FirTables& ft) {
  const int m = fir_fft_size(n);
  int ns = 0, r[4] = {0, 0, 0, 0};
  const bool mono = plan_of_m(m, ns, r);
  // n = 8192 (m = 16384) has no monolithic plan: it runs on the split / circular forms only
  if (!mono && !(m == 2 * n && n == 8192))
    return set_err(c, THZ_EINVAL, "trace too long for the FIR transform (n must be <= 7943 or exactly 8192)");
  ft.mono = mono;
  uint64_t key = 0xcbf29ce484222325ull;
  for (int b = 0; b < B; ++b) key = fnv1a(bands[b].fir, sizeof(bands[b].fir), key);
  key = fnv1a(&m, sizeof m, key);
  void* dp = nullptr;
  // layout (floats): hq [B][m] | wq [B][m/2] | wnyq [B, padded to 4] | edge [2][B][512] float2
  const size_t n_hq = mono ? (size_t)B * m : 0, n_wq = mono ? (size_t)B * (m / 2) : 0, n_ny = (size_t)((B + 3) & ~3);
  const size_t n_edge = (size_t)2 * B * 512 * 2;
  const int nh = m / 2;
  int nsh, rh[4];
  // split / circular forms: N-point geometry (two CTAs per SM up to n = 4096, one 512-thread CTA at 8192)
  const bool split = (nh == n) && nh >= 256 && nh <= 8192 && plan_of_m(nh, nsh, rh);
  const int Bp = (B + 3) & ~3;
  const size_t n_split_base = split ? ((size_t)2 * B * nh + (size_t)2 * B * (nh / 2) + (size_t)2 * nh) : 0;
  const size_t n_il = split ? ((size_t)3 * (nh / 2) * Bp + Bp) : 0;   // we4, wo4, he4, hny4
  const size_t n_split = n_split_base + n_il;
  int rc = ws_get(c, WS_FIR, (n_hq + n_wq + n_ny + n_edge + n_split) * sizeof(float), &dp);
  if (rc != THZ_OK) return rc;
  ft.d_hq = (float*)dp;
  ft.d_wq = ft.d_hq + n_hq;
  ft.d_wnyq = ft.d_wq + n_wq;
  ft.d_edge = reinterpret_cast<float2*>(ft.d_wnyq + n_ny);
  ft.split = split;
  if (split) {
    ft.d_he = ft.d_wnyq + n_ny + n_edge;
    ft.d_ho = ft.d_he + (size_t)B * nh;
    ft.d_we = ft.d_ho + (size_t)B * nh;
    ft.d_wo = ft.d_we + (size_t)B * (nh / 2);
    ft.d_mod = reinterpret_cast<float2*>(ft.d_wo + (size_t)B * (nh / 2));
    ft.Bp = Bp;
    ft.d_we4 = ft.d_he + n_split_base;
    ft.d_wo4 = ft.d_we4 + (size_t)(nh / 2) * Bp;
    ft.d_he4 = ft.d_wo4 + (size_t)(nh / 2) * Bp;
    ft.d_hny4 = ft.d_he4 + (size_t)(nh / 2) * Bp;
  }
  ft.m = m;
  ft.B = B;
  if (c->fir_key == key && c->fir_m == m) return THZ_OK;
  std::vector<float> all(n_hq + n_wq + n_ny + n_edge + n_split, 0.f), one;
  const int RLs = mono ? r[ns - 1] : 1;
  for (int b = 0; b < B; ++b) {
    std::vector<float> hnat;
    fir_spectrum_natural(m, bands[b].fir, hnat);
    if (mono) {
      if (build_fir_hq(m, bands[b].fir, one) != THZ_OK) return set_err(c, THZ_EINVAL, "bad FIR transform size");
      std::copy(one.begin(), one.end(), all.begin() + (size_t)b * m);
      // Parseval weights |H|^2 / m = (H/m)^2 * m for the lower-half registers (last-stage digit < RL/2):
      // they are the first m/2 entries of the [digit][beta] register-order table
      for (int i = 0; i < m / 2; ++i) all[n_hq + (size_t)b * (m / 2) + i] = one[i] * one[i] * (float)m;
    }
    // Nyquist bin of the m-point spectrum
    const float hn = hnat[m / 2];
    all[n_hq + n_wq + b] = hn * hn * (float)m;
    std::vector<float2> head, tail;
    build_edge_tables(bands[b].fir, head, tail);
    float* eh = all.data() + n_hq + n_wq + n_ny + ((size_t)0 * B + b) * 512 * 2;
    float* et = all.data() + n_hq + n_wq + n_ny + ((size_t)1 * B + b) * 512 * 2;
    memcpy(eh, head.data(), 512 * sizeof(float2));
    memcpy(et, tail.data(), 512 * sizeof(float2));
    if (split) {
      // the even / odd sub-spectra of the natural-order H / m, in the register order of the (m/2)-point plan
      const int RLh = rh[nsh - 1];
      float* he = all.data() + n_hq + n_wq + n_ny + n_edge + (size_t)b * nh;
      float* ho = he + (size_t)B * nh;
      float* we = all.data() + n_hq + n_wq + n_ny + n_edge + (size_t)2 * B * nh + (size_t)b * (nh / 2);
      float* wo = we + (size_t)B * (nh / 2);
      for (int beta = 0; beta < nh / RLh; ++beta)
        for (int mm = 0; mm < RLh; ++mm) {
          int pp = beta * RLh + mm, j = 0, w = 1, L = nh;
          for (int s2 = 0; s2 < nsh; ++s2) {
            const int S = L / rh[s2];
            const int q = pp / S;
            pp -= q * S;
            j += q * w;
            w *= rh[s2];
            L = S;
          }
          const size_t o = (size_t)mm * (nh / RLh) + beta;
          he[o] = hnat[2 * j];
          ho[o] = hnat[2 * j + 1];
          if (o < (size_t)(nh / 2)) {   // lower-half registers
            we[o] = he[o] * he[o] * (float)m;
            wo[o] = ho[o] * ho[o] * (float)m;
          }
        }
      float* il = all.data() + n_hq + n_wq + n_ny + n_edge + n_split_base;
      float* we4 = il;
      float* wo4 = we4 + (size_t)(nh / 2) * Bp;
      float* he4 = wo4 + (size_t)(nh / 2) * Bp;
      float* hny4 = he4 + (size_t)(nh / 2) * Bp;
      for (int o = 0; o < nh / 2; ++o) {   // [band group of 4][register-order index][band within the group]
        const size_t q = ((size_t)(b >> 2) * (nh / 2) + o) * 4 + (b & 3);
        we4[q] = we[o];
        wo4[q] = wo[o];
        he4[q] = he[o];
      }
      hny4[b] = he[(size_t)(RLh / 2) * (nh / RLh)];
    }
  }
  if (split) {
    float2* mod = reinterpret_cast<float2*>(all.data() + n_hq + n_wq + n_ny + n_edge + (size_t)2 * B * nh +
                                            (size_t)2 * B * (nh / 2));
    for (int i = 0; i < nh; ++i) {
      const double ang = -2.0 * M_PI * (double)i / (double)m;
      mod[i] = make_float2((float)cos(ang), (float)sin(ang));
    }
  }
  THZ_CUDA(c, cudaStreamSynchronize(s));   // no kernel still reads the previous spectra
  THZ_CUDA(c, cudaMemcpyAsync(ft.d_hq, all.data(), all.size() * sizeof(float), cudaMemcpyHostToDevice, s));
  THZ_CUDA(c, cudaStreamSynchronize(s));
  c->fir_key = key;
  c->fir_m = m;
  return THZ_OK;
}

template <int M, typename K>
static int launch_fir(thz_ctx* c, cudaStream_t s, K kernel, const FirArgs& a, size_t smem_override = 0,
                      int carveout_pct = -1) {
  using GEO = DGeo<M>;
  const size_t smem = smem_override ? smem_override : GEO::smem_bytes;
  const void* key = (const void*)kernel;
  auto it = c->occ.find(key);
  if (it == c->occ.end()) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaFuncSetAttribute(smem)");
    if (carveout_pct >= 0) {   // leave the rest of the unified L1 / shared memory to the table loads
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_pct);
      if (e != cudaSuccess) return cuda_fail(c, e, "cudaFuncSetAttribute(carveout)");
    }
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, GEO::NT, smem);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (nb < 1) return set_err(c, THZ_ECUDA, "FIR kernel does not fit on an SM");
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
  if (e != cudaSuccess) return cuda_fail(c, e, "FIR kernel launch");
  return THZ_OK;
}

template <int M> static int do_energy(thz_ctx* c, cudaStream_t s, const FirArgs& a) {
  return launch_fir<M>(c, s, k_fir_energy<M>, a);
}
template <int M> static int do_apply(thz_ctx* c, cudaStream_t s, const FirArgs& a) {
  return launch_fir<M>(c, s, k_fir_apply<M>, a);
}
template <int M> static int do_energy_total(thz_ctx* c, cudaStream_t s, const FirArgs& a) {
  return launch_fir<M>(c, s, k_fir_energy_total<M>, a);
}
// split kernels use the N-point geometry (identical to DGeo<N>)
template <int N> static int do_energy_split(thz_ctx* c, cudaStream_t s, const FirArgs& a) {
  if (c->unstaged_fir) return launch_fir<N>(c, s, k_fir_energy_split<N, false>, a, SGeo<N>::base_bytes, 40);
  return launch_fir<N>(c, s, k_fir_energy_split<N, true>, a, SGeo<N>::smem_bytes);
}
template <int N> static int do_apply_split(thz_ctx* c, cudaStream_t s, const FirArgs& a) {
  return launch_fir<N>(c, s, k_fir_apply_split<N>, a, SGeo<N>::smem_bytes);
}
template <int N> static int do_apply_circ(thz_ctx* c, cudaStream_t s, const FirArgs& a) {
  if (a.xspec != nullptr) return launch_fir<N>(c, s, k_fir_apply_circ<N, false, true>, a, SGeo<N>::base_bytes, 40);
  if (c->unstaged_fir) return launch_fir<N>(c, s, k_fir_apply_circ<N, false>, a, SGeo<N>::base_bytes, 40);
  return launch_fir<N>(c, s, k_fir_apply_circ<N, true>, a, SGeo<N>::smem_bytes);
}

#define THZ_DISPATCH_M(m, FN, ...)                 \
  switch (m) {                                     \
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

static int dispatch_energy(thz_ctx* c, cudaStream_t s, int m, const FirArgs& a) { THZ_DISPATCH_M(m, do_energy, c, s, a); }
static int dispatch_apply(thz_ctx* c, cudaStream_t s, int m, const FirArgs& a) { THZ_DISPATCH_M(m, do_apply, c, s, a); }
static int dispatch_energy_total(thz_ctx* c, cudaStream_t s, int m, const FirArgs& a) {
  THZ_DISPATCH_M(m, do_energy_total, c, s, a);
}
static int dispatch_energy_split(thz_ctx* c, cudaStream_t s, int n, const FirArgs& a) {
  THZ_DISPATCH_M(n, do_energy_split, c, s, a);
}
static int dispatch_apply_split(thz_ctx* c, cudaStream_t s, int n, const FirArgs& a) {
  THZ_DISPATCH_M(n, do_apply_split, c, s, a);
}
static int dispatch_apply_circ(thz_ctx* c, cudaStream_t s, int n, const FirArgs& a) {
  THZ_DISPATCH_M(n, do_apply_circ, c, s, a);
}

// event pair around one cube kernel while thz_deconvolution_dev collects its per-kernel breakdown; the pairs
// are resolved after the final stream synchronisation, the launches themselves stay asynchronous
struct KernelTimer {
  thz_ctx* c;
  cudaStream_t s;
  int slot;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  KernelTimer(thz_ctx* c_, cudaStream_t s_, int slot_) : c(c_), s(s_), slot(slot_) {
    if (!c->time_kernels) return;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e0 = nullptr; return; }
    cudaEventRecord(e0, s);
  }
  void stop() {
    if (!e0) return;
    cudaEventRecord(e1, s);
    c->kernel_events.push_back({slot, e0, e1});
  }
};

static void resolve_kernel_events(thz_ctx* c) {
  for (auto& ev : c->kernel_events) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev.e0, ev.e1) == cudaSuccess) c->kernel_ms[ev.slot] += ms;
    cudaEventDestroy(ev.e0);
    cudaEventDestroy(ev.e1);
  }
  c->kernel_events.clear();
}

int deconv_energies(thz_ctx* c, cudaStream_t s, const float* d_cube, int64_t P, int n, const thz_band_plan* bands,
                    int B, float* d_energy, int64_t bstride) {
  if (bstride == 0) bstride = P;
  if (B < 1 || B > THZ_MAX_BANDS || !bands) return set_err(c, THZ_EINVAL, "bad band count");
  if (P == 0) return THZ_OK;
  if (!d_cube || !d_energy || n < 2) return set_err(c, THZ_EINVAL, "null pointer");
  if ((reinterpret_cast<uintptr_t>(d_cube) & 15u) != 0) return set_err(c, THZ_EINVAL, "cube must be 16-byte aligned");
  FirTables ft;
  int rc = upload_fir_tables(c, s, n, bands, B, ft);
  if (rc != THZ_OK) return rc;
  const FftTables* tb = nullptr;
  if (ft.mono) rc = get_tables(c, ft.m, &tb);
  if (rc == THZ_OK) {
    FirArgs a{};
    a.x = d_cube; a.n = n; a.P = P; a.hq = ft.d_hq; a.B = B; a.energy = d_energy; a.tw = tb ? tb->d_tw : nullptr;
    a.bstride = bstride;
    a.wq = ft.d_wq; a.wnyq = ft.d_wnyq; a.edge = ft.d_edge;
    if (n >= 512 && ft.m >= n + THZ_FIR_TAPS - 1) {
      // Parseval total energy of the full linear convolution minus the two excluded edge segments
      const FftTables* tb512 = nullptr;
      rc = get_tables(c, 512, &tb512);
      if (rc != THZ_OK) return rc;
      a.tw512 = tb512->d_tw;
      if (ft.split) {
        const FftTables* tbn = nullptr;
        rc = get_tables(c, n, &tbn);
        if (rc != THZ_OK) return rc;
        FirArgs as = a;
        as.tw = tbn->d_tw;
        as.he = ft.d_he; as.ho = ft.d_ho; as.we = ft.d_we; as.wo = ft.d_wo; as.mod = ft.d_mod;
        as.Bp = ft.Bp; as.we4 = ft.d_we4; as.wo4 = ft.d_wo4;
        KernelTimer kt(c, s, 0);
        rc = dispatch_energy_split(c, s, n, as);
        kt.stop();
      } else {
        KernelTimer kt(c, s, 0);
        rc = dispatch_energy_total(c, s, ft.m, a);
        kt.stop();
      }
      if (rc == THZ_OK) {
        KernelTimer kt(c, s, 1);
        if (c->edge_mma && edges_mma_supported(n))   // Toeplitz GEMM on the tensor cores (thz_edges_mma.cu)
          rc = launch_fir_edges_mma(c, s, d_cube, P, n, bands, B, d_energy, bstride);
        else
          rc = launch_fir<512>(c, s, k_fir_edges, a);
        kt.stop();
      }
    } else {
      rc = dispatch_energy(c, s, ft.m, a);   // short traces: B inverse transforms per pair
    }
  }
  return rc;
}

// Slots 2..7 of the chain on d_in -> d_out (+ intensity) and the band energies of the filtered traces.  One fused
// kernel + the edge kernel when the plan is a power-of-two length >= 512 with the split FIR form (every BASELINE
// shape); the trace pass followed by deconv_energies otherwise, or with THZ_CHAIN_FUSE=off (A/B checks).
// c->chain_even: how the fused kernel obtains the even bins -- 0 = from the filtered spectrum when the gate allows
// it (default), 1 = always by a forward transform of the stored trace (THZ_CHAIN_EVEN=transform, A/B checks).
bool chain_spectral_ok(const thz_ctx* c, int n, int64_t P) {
  return c->chain_fuse && c->chain_spectral && !c->force_split_apply && c->plan.blue_m == 0 && c->plan.n == n &&
         n >= 512 && supported_n(n) && (P % 2) == 0;
}

int chain_energies(thz_ctx* c, cudaStream_t s, const float* d_in, float* d_out, float* d_img, int64_t P, int n,
                   const thz_band_plan* bands, int B, float* d_energy, int64_t bstride, float* d_edges) {
  if (bstride == 0) bstride = P;
  if (d_edges && !chain_spectral_ok(c, n, P)) return set_err(c, THZ_ESTATE, "spectral hand-off not available for this plan");
  if (c->plan.n != n) return set_err(c, THZ_ESTATE, "thz_plan_trace(n, ...) must be called first");
  if (B < 1 || B > THZ_MAX_BANDS || !bands) return set_err(c, THZ_EINVAL, "bad band count");
  if (P == 0) return THZ_OK;
  if (!d_in || !d_out || !d_energy) return set_err(c, THZ_EINVAL, "null pointer");
  bool fuse = c->chain_fuse && c->plan.blue_m == 0 && n >= 512 && supported_n(n);
  FirTables ft;
  int rc = THZ_OK;
  if (fuse) {
    rc = upload_fir_tables(c, s, n, bands, B, ft);
    if (rc != THZ_OK) return rc;
    fuse = ft.split && ft.m >= n + THZ_FIR_TAPS - 1;
  }
  if (!fuse) {
    if (d_edges) return set_err(c, THZ_ESTATE, "spectral hand-off needs the fused kernel");
    KernelTimer kt(c, s, 4);
    rc = launch_trace_fused(c, s, d_in, d_out, d_img, P);
    kt.stop();
    if (rc == THZ_OK) rc = deconv_energies(c, s, d_out, P, n, bands, B, d_energy, bstride);
    return rc;
  }
  TraceArgs ta;
  rc = trace_fused_args(c, ta, d_in, d_out, d_img, P);
  if (rc != THZ_OK) return rc;
  const FftTables *tbn = nullptr, *tb512 = nullptr;
  rc = get_tables(c, n, &tbn);
  if (rc == THZ_OK) rc = get_tables(c, 512, &tb512);
  if (rc != THZ_OK) return rc;
  FirArgs a{};
  a.x = d_out; a.n = n; a.P = P; a.hq = ft.d_hq; a.B = B; a.energy = d_energy; a.bstride = bstride;
  a.wq = ft.d_wq; a.wnyq = ft.d_wnyq; a.edge = ft.d_edge; a.tw512 = tb512->d_tw;
  FirArgs as = a;
  as.tw = tbn->d_tw;
  as.he = ft.d_he; as.ho = ft.d_ho; as.we = ft.d_we; as.wo = ft.d_wo; as.mod = ft.d_mod;
  as.Bp = ft.Bp; as.we4 = ft.d_we4; as.wo4 = ft.d_wo4;
  as.edges = d_edges;
  if (d_edges) {   // the edge kernels read the side buffer as a cube of 512-sample rows
    a.x = d_edges;
    a.n = 512;
  }
  const int post = c->chain_even_transform ? 2 : c->plan.post_mode;
  {
    KernelTimer kt(c, s, 0);
    rc = dispatch_chain_fused(c, s, n, ta, as, post);
    kt.stop();
  }
  if (rc == THZ_OK) {
    KernelTimer kt(c, s, 1);
    if (c->edge_mma && edges_mma_supported(n))
      rc = d_edges ? launch_fir_edges_mma(c, s, d_edges, P, 512, bands, B, d_energy, bstride, true)
                   : launch_fir_edges_mma(c, s, d_out, P, n, bands, B, d_energy, bstride);
    else
      rc = launch_fir<512>(c, s, k_fir_edges, a);
    kt.stop();
  }
  return rc;
}

// `lane` selects the wrap-around correction workspace: calls that are in flight on different streams at the same
// time (the chunk pipeline of thz_chain_host) must not share one (k_fir_edge_corr of chunk i+1 would overwrite
// what k_fir_apply_circ of chunk i still reads).  Lane 0 = the context's compute stream, 1 + k = hstream[k].
int deconv_apply(thz_ctx* c, cudaStream_t s, const float* d_cube, const float* d_gain, int64_t P, int n,
                 const thz_band_plan* bands, int B, float* d_out, float* d_img, int64_t bstride, int lane,
                 const float* d_edges) {
  if (bstride == 0) bstride = P;
  if (d_edges && !chain_spectral_ok(c, n, P)) return set_err(c, THZ_ESTATE, "spectral hand-off not available for this plan");
  if (B < 1 || B > THZ_MAX_BANDS || !bands) return set_err(c, THZ_EINVAL, "bad band count");
  if (P == 0) return THZ_OK;
  if (!d_cube || !d_gain || !d_out || n < 2) return set_err(c, THZ_EINVAL, "null pointer");
  if ((reinterpret_cast<uintptr_t>(d_cube) & 15u) != 0) return set_err(c, THZ_EINVAL, "cube must be 16-byte aligned");
  FirTables ft;
  int rc = upload_fir_tables(c, s, n, bands, B, ft);
  if (rc != THZ_OK) return rc;
  const FftTables* tb = nullptr;
  if (ft.mono) rc = get_tables(c, ft.m, &tb);
  if (rc == THZ_OK) {
    FirArgs a{};
    a.x = d_cube; a.n = n; a.P = P; a.hq = ft.d_hq; a.B = B; a.gain = d_gain; a.out = d_out; a.img = d_img;
    a.bstride = bstride;
    a.tw = tb ? tb->d_tw : nullptr;
    if (ft.split) {
      const FftTables* tbn = nullptr;
      rc = get_tables(c, n, &tbn);
      if (rc != THZ_OK) return rc;
      a.tw = tbn->d_tw;
      a.he = ft.d_he; a.ho = ft.d_ho; a.mod = ft.d_mod;
      a.Bp = ft.Bp; a.he4 = ft.d_he4; a.hny4 = ft.d_hny4;
      if (n >= 512 && !c->force_split_apply) {
        // circular form: edge corrections of a chunk of pairs, then the one-transform-pair main pass
        const FftTables* tb512 = nullptr;
        rc = get_tables(c, 512, &tb512);
        if (rc != THZ_OK) return rc;
        a.tw512 = tb512->d_tw;
        a.edge = ft.d_edge;
        const int64_t npairs = (P + 1) / 2;
        const int64_t chunk_pairs = std::min<int64_t>(npairs, kCorrChunkPairs);
        void* pc = nullptr;
        rc = ws_get(c, WS_EDGE_CORR + lane, (size_t)chunk_pairs * 2 * kCorrStride * sizeof(float2), &pc);
        if (rc != THZ_OK) return rc;
        for (int64_t q0 = 0; rc == THZ_OK && q0 < npairs; q0 += chunk_pairs) {
          const int64_t p_lo = 2 * q0, p_hi = std::min<int64_t>(P, 2 * (q0 + chunk_pairs));
          FirArgs ac = a;
          ac.x = d_cube + p_lo * n;
          ac.P = p_hi - p_lo;
          ac.gain = d_gain + p_lo;
          ac.out = d_out + p_lo * n;
          ac.img = d_img ? d_img + p_lo : nullptr;
          ac.corr = (float2*)pc;
          {
            FirArgs ae = ac;
            if (d_edges) {   // spectral hand-off: edge samples from the side buffer (rows of 512 floats)
              ae.x = d_edges + p_lo * 512;
              ae.n = 512;
            }
            KernelTimer kt(c, s, 2);
            rc = launch_fir<512>(c, s, k_fir_edge_corr, ae);
            kt.stop();
          }
          if (d_edges) ac.xspec = reinterpret_cast<const float2*>(d_cube) + (p_lo / 2) * n;
          if (rc == THZ_OK) {
            KernelTimer kt(c, s, 3);
            rc = dispatch_apply_circ(c, s, n, ac);
            kt.stop();
          }
        }
      } else {
        KernelTimer kt(c, s, 3);
        rc = dispatch_apply_split(c, s, n, a);
        kt.stop();
      }
    } else {
      KernelTimer kt(c, s, 3);
      rc = dispatch_apply(c, s, ft.m, a);
      kt.stop();
    }
  }
  return rc;
}


// ---- the two cube-touching halves of the host-pointer chain (thz_chain_host, thz_group_chain_host) ----
// in : H2D chunks -> fused trace pass -> band energies; the filtered cube stays in WS_HOST_CUBE, the intensity in
//      WS_HOST_IMG, the energies in WS_ENERGY [n_bands][P].  Without bands the filtered chunks go straight back.
int chain_pass_in(thz_ctx* c, const float* cube, int64_t P, int n, const thz_band_plan* bands, int n_bands, float* out,
                  float** d_energy_out, float** d_gain_out) {
  if (c->plan.n != n) return set_err(c, THZ_ESTATE, "thz_plan_trace(n, ...) must be called first");
  if (!cube || !out) return set_err(c, THZ_EINVAL, "null pointer");
  if (n_bands < 0 || n_bands > THZ_MAX_BANDS || (n_bands > 0 && !bands)) return set_err(c, THZ_EINVAL, "bad bands");
  void *pc = nullptr, *pi = nullptr, *pe = nullptr, *pg = nullptr;
  int rc = ws_get(c, WS_HOST_CUBE, (size_t)P * n * sizeof(float), &pc);
  if (rc == THZ_OK) rc = ws_get(c, WS_HOST_IMG, (size_t)P * sizeof(float), &pi);
  if (rc == THZ_OK && n_bands) rc = ws_get(c, WS_ENERGY, (size_t)n_bands * P * sizeof(float), &pe);
  if (rc == THZ_OK && n_bands) rc = ws_get(c, WS_GAIN, (size_t)n_bands * P * sizeof(float), &pg);
  if (rc != THZ_OK) return rc;
  float *d_cube = (float*)pc, *d_img = (float*)pi, *d_energy = (float*)pe;
  if (d_energy_out) *d_energy_out = d_energy;
  if (d_gain_out) *d_gain_out = (float*)pg;
  // FIR tables are built (and cached) before the pipelined loop so that no chunk waits on the host
  float* d_edges = nullptr;
  c->host_chain_spectral = false;
  if (n_bands) {
    FirTables ft;
    rc = upload_fir_tables(c, c->stream, n, bands, n_bands, ft);
    if (rc != THZ_OK) return rc;
    if (chain_spectral_ok(c, n, P)) {   // the resident cube holds spectra between the two halves (see k_chain_energy_fused)
      void* pr = nullptr;
      rc = ws_get(c, WS_EDGE_ROWS, (size_t)P * 512 * sizeof(float), &pr);
      if (rc != THZ_OK) return rc;
      d_edges = (float*)pr;
      c->host_chain_spectral = true;
    }
  }
  int64_t ct = (int64_t)c->host_chunk_bytes / ((int64_t)n * 4);   // 256 MiB chunks (THZ_CHAIN_CHUNK_BYTES), whole pairs
  ct &= ~(int64_t)1;
  if (ct < 2) ct = 2;
  int i = 0;
  for (int64_t p = 0; p < P && rc == THZ_OK; p += ct, ++i) {
    const int64_t np = std::min(ct, P - p);
    cudaStream_t s = c->hstream[i % kHostStreams];
    float* d = d_cube + p * n;
    THZ_CUDA(c, cudaMemcpyAsync(d, cube + p * n, (size_t)np * n * sizeof(float), cudaMemcpyHostToDevice, s));
    if (n_bands) rc = chain_energies(c, s, d, d, d_img + p, np, n, bands, n_bands, d_energy + p, P,
                                     d_edges ? d_edges + p * 512 : nullptr);
    else rc = launch_trace_fused(c, s, d, d, d_img + p, np);
    if (rc == THZ_OK && !n_bands) {
      THZ_CUDA(c, cudaMemcpyAsync(out + p * n, d, (size_t)np * n * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
  }
  for (int k = 0; k < kHostStreams; ++k) THZ_CUDA(c, cudaStreamSynchronize(c->hstream[k]));
  return rc;
}

// out: gain application chunks overlapped with their D2H copies, then the intensity image
int chain_pass_out(thz_ctx* c, int64_t P, int n, const thz_band_plan* bands, int n_bands, float* out, float* img) {
  float* d_cube = (float*)c->ws[WS_HOST_CUBE].first;
  float* d_img = (float*)c->ws[WS_HOST_IMG].first;
  float* d_gain = n_bands ? (float*)c->ws[WS_GAIN].first : nullptr;
  if (!d_cube || !d_img || (n_bands && !d_gain)) return set_err(c, THZ_ESTATE, "chain_pass_in has not run");
  const float* d_edges = (n_bands && c->host_chain_spectral) ? (const float*)c->ws[WS_EDGE_ROWS].first : nullptr;
  int rc = THZ_OK;
  if (n_bands) {
    int64_t ct = (int64_t)c->host_chunk_bytes / ((int64_t)n * 4);
    ct &= ~(int64_t)1;
    if (ct < 2) ct = 2;
    int i = 0;
    for (int64_t p = 0; p < P && rc == THZ_OK; p += ct, ++i) {
      const int64_t np = std::min(ct, P - p);
      cudaStream_t s = c->hstream[i % kHostStreams];
      float* d = d_cube + p * n;
      rc = deconv_apply(c, s, d, d_gain + p, np, n, bands, n_bands, d, d_img + p, P, 1 + i % kHostStreams,
                        d_edges ? d_edges + p * 512 : nullptr);
      if (rc == THZ_OK)
        THZ_CUDA(c, cudaMemcpyAsync(out + p * n, d, (size_t)np * n * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
    for (int k = 0; k < kHostStreams; ++k) THZ_CUDA(c, cudaStreamSynchronize(c->hstream[k]));
    if (rc != THZ_OK) return rc;
  }
  if (img) {
    THZ_CUDA(c, cudaMemcpyAsync(img, d_img, (size_t)P * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return THZ_OK;
}

}  // namespace thz

using namespace thz;

#define CHECK_CTX(c)                                                   \
  do {                                                                 \
    if (!(c)) return THZ_EINVAL;                                       \
    cudaError_t e_ = cudaSetDevice((c)->device);                       \
    if (e_ != cudaSuccess) return cuda_fail((c), e_, "cudaSetDevice"); \
  } while (0)

// Richardson-Lucy of all band images on this GPU: iteration i of every band that still iterates batched into one
// launch per filtering (richardson_lucy_bands); band after band when a PSF does not fit that kernel or when
// THZ_RL_BATCH=off (A/B checks).
static int rl_all_bands(thz_ctx* c, const float* d_energy, int64_t P, int rows, int cols, const thz_band_plan* bands,
                        int n_bands, float* d_gain, const volatile uint8_t* abort_flag, thz_progress_fn progress,
                        void* progress_user) {
  if (abort_flag && *abort_flag) return THZ_ABORTED;
  int rc = THZ_SKIP_NO_PSF;
  if (c->rl_batch)
    rc = richardson_lucy_bands(c, c->stream, d_energy, P, rows, cols, bands, n_bands, d_gain, abort_flag, progress,
                               progress_user, nullptr);
  if (rc != THZ_SKIP_NO_PSF) return rc;
  long total_iter = 0, done_iter = 0;
  for (int b = 0; b < n_bands; ++b) total_iter += std::max(bands[b].n_iter, 1);
  rc = THZ_OK;
  for (int b = 0; rc == THZ_OK && b < n_bands; ++b) {
    if (abort_flag && *abort_flag) return THZ_ABORTED;
    const float base = 0.1f + 0.8f * (float)done_iter / (float)total_iter;
    const float span = 0.8f * (float)std::max(bands[b].n_iter, 1) / (float)total_iter;
    rc = richardson_lucy(c, c->stream, d_energy + (size_t)b * P, rows, cols, bands[b].psf_x, bands[b].kx,
                         bands[b].psf_y, bands[b].ky, nullptr, bands[b].direct, bands[b].n_iter, nullptr,
                         d_gain + (size_t)b * P, abort_flag, progress, progress_user, base, span);
    done_iter += std::max(bands[b].n_iter, 1);
  }
  return rc;
}

extern "C" {

int thz_deconv_energies_dev(thz_ctx* c, const float* d_cube, int64_t P, int n, const thz_band_plan* bands, int n_bands,
                            float* d_energy) {
  CHECK_CTX(c);
  return deconv_energies(c, c->stream, d_cube, P, n, bands, n_bands, d_energy);
}

int thz_deconv_apply_dev(thz_ctx* c, const float* d_cube, const float* d_gain, int64_t P, int n,
                         const thz_band_plan* bands, int n_bands, float* d_out, float* d_img) {
  CHECK_CTX(c);
  return deconv_apply(c, c->stream, d_cube, d_gain, P, n, bands, n_bands, d_out, d_img);
}

int thz_deconvolution_dev(thz_ctx* c, const float* d_cube, int rows, int cols, int n, const thz_band_plan* bands,
                          int n_bands, float* d_out, float* d_img, const volatile uint8_t* abort_flag,
                          thz_progress_fn progress, void* progress_user) {
  CHECK_CTX(c);
  if (!bands || n_bands < 1 || n_bands > THZ_MAX_BANDS) return set_err(c, THZ_EINVAL, "bad band count");
  const int64_t P = (int64_t)rows * cols;
  if (P == 0) return THZ_OK;
  if (progress) progress(0.0f, progress_user);
  void *pe = nullptr, *pg = nullptr;
  int rcw = ws_get(c, WS_ENERGY, (size_t)n_bands * P * sizeof(float), &pe);
  if (rcw == THZ_OK) rcw = ws_get(c, WS_GAIN, (size_t)n_bands * P * sizeof(float), &pg);
  if (rcw != THZ_OK) return rcw;
  float *d_energy = (float*)pe, *d_gain = (float*)pg;
  cudaEvent_t ev[4];
  for (auto& e : ev) cudaEventCreate(&e);
  for (float& v : c->kernel_ms) v = 0.f;
  c->time_kernels = true;
  cudaEventRecord(ev[0], c->stream);
  int rc = deconv_energies(c, c->stream, d_cube, P, n, bands, n_bands, d_energy);
  cudaEventRecord(ev[1], c->stream);
  long total_iter = 0, done_iter = 0;
  for (int b = 0; b < n_bands; ++b) total_iter += std::max(bands[b].n_iter, 1);
  if (rc == THZ_OK) rc = rl_all_bands(c, d_energy, P, rows, cols, bands, n_bands, d_gain, abort_flag, progress, progress_user);
  done_iter = total_iter;
  cudaEventRecord(ev[2], c->stream);
  if (rc == THZ_OK) rc = deconv_apply(c, c->stream, d_cube, d_gain, P, n, bands, n_bands, d_out, d_img);
  cudaEventRecord(ev[3], c->stream);
  c->time_kernels = false;
  cudaStreamSynchronize(c->stream);
  resolve_kernel_events(c);
  for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&c->stage_ms[i], ev[i], ev[i + 1]);
  c->stage_ms[3] = (float)done_iter;
  for (auto& e : ev) cudaEventDestroy(e);
  if (rc == THZ_OK && progress) progress(1.0f, progress_user);
  return rc;
}

int thz_chain_energies_dev(thz_ctx* c, const float* d_in, float* d_out, float* d_img, int64_t P, int n,
                           const thz_band_plan* bands, int n_bands, float* d_energy) {
  CHECK_CTX(c);
  return chain_energies(c, c->stream, d_in, d_out, d_img, P, n, bands, n_bands, d_energy);
}

// The whole default chain + deconvolution on a device-resident cube (the device-pointer twin of thz_chain_host):
// trace pass fused with the band energies -> Richardson-Lucy -> gain application.  d_out may alias d_in.
int thz_chain_dev(thz_ctx* c, const float* d_in, int rows, int cols, int n, const thz_band_plan* bands, int n_bands,
                  float* d_out, float* d_img, const volatile uint8_t* abort_flag, thz_progress_fn progress,
                  void* progress_user) {
  CHECK_CTX(c);
  if (!bands || n_bands < 1 || n_bands > THZ_MAX_BANDS) return set_err(c, THZ_EINVAL, "bad band count");
  const int64_t P = (int64_t)rows * cols;
  if (P == 0) return THZ_OK;
  if (!d_img) return set_err(c, THZ_EINVAL, "null intensity pointer");
  if (progress) progress(0.0f, progress_user);
  void *pe = nullptr, *pg = nullptr;
  int rcw = ws_get(c, WS_ENERGY, (size_t)n_bands * P * sizeof(float), &pe);
  if (rcw == THZ_OK) rcw = ws_get(c, WS_GAIN, (size_t)n_bands * P * sizeof(float), &pg);
  if (rcw != THZ_OK) return rcw;
  float *d_energy = (float*)pe, *d_gain = (float*)pg;
  float* d_edges = nullptr;
  if (chain_spectral_ok(c, n, P)) {
    void* pr = nullptr;
    rcw = ws_get(c, WS_EDGE_ROWS, (size_t)P * 512 * sizeof(float), &pr);
    if (rcw != THZ_OK) return rcw;
    d_edges = (float*)pr;
  }
  cudaEvent_t ev[4];
  for (auto& e : ev) cudaEventCreate(&e);
  resolve_kernel_events(c);
  for (float& v : c->kernel_ms) v = 0.f;
  c->time_kernels = true;
  cudaEventRecord(ev[0], c->stream);
  int rc = chain_energies(c, c->stream, d_in, d_out, d_img, P, n, bands, n_bands, d_energy, 0, d_edges);
  cudaEventRecord(ev[1], c->stream);
  long total_iter = 0;
  for (int b = 0; b < n_bands; ++b) total_iter += std::max(bands[b].n_iter, 1);
  if (rc == THZ_OK) rc = rl_all_bands(c, d_energy, P, rows, cols, bands, n_bands, d_gain, abort_flag, progress, progress_user);
  cudaEventRecord(ev[2], c->stream);
  if (rc == THZ_OK) rc = deconv_apply(c, c->stream, d_out, d_gain, P, n, bands, n_bands, d_out, d_img, 0, 0, d_edges);
  cudaEventRecord(ev[3], c->stream);
  c->time_kernels = false;
  cudaStreamSynchronize(c->stream);
  resolve_kernel_events(c);
  for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&c->stage_ms[i], ev[i], ev[i + 1]);
  c->stage_ms[3] = (float)total_iter;
  for (auto& e : ev) cudaEventDestroy(e);
  if (rc == THZ_OK && progress) progress(1.0f, progress_user);
  return rc;
}

// thz_chain_dev in two calls for a host that runs Richardson-Lucy itself between them (one rank of a multi-GPU run:
// thz_slab_rl).  d_work [P][n] holds the hand-off between the halves in a private layout (the spectra of the
// filtered pairs when the plan allows the spectral hand-off, the filtered traces otherwise) and must not be
// touched in between; d_out of the second half may alias it.
int thz_chain_begin_dev(thz_ctx* c, const float* d_in, float* d_work, float* d_img, int64_t P, int n,
                        const thz_band_plan* bands, int n_bands, float* d_energy) {
  CHECK_CTX(c);
  float* d_edges = nullptr;
  c->host_chain_spectral = false;
  if (P > 0 && chain_spectral_ok(c, n, P)) {
    void* pr = nullptr;
    int rc = ws_get(c, WS_EDGE_ROWS, (size_t)P * 512 * sizeof(float), &pr);
    if (rc != THZ_OK) return rc;
    d_edges = (float*)pr;
    c->host_chain_spectral = true;
  }
  return chain_energies(c, c->stream, d_in, d_work, d_img, P, n, bands, n_bands, d_energy, 0, d_edges);
}

int thz_chain_end_dev(thz_ctx* c, const float* d_work, const float* d_gain, int64_t P, int n, const thz_band_plan* bands,
                      int n_bands, float* d_out, float* d_img) {
  CHECK_CTX(c);
  const float* d_edges = c->host_chain_spectral ? (const float*)c->ws[WS_EDGE_ROWS].first : nullptr;
  return deconv_apply(c, c->stream, d_work, d_gain, P, n, bands, n_bands, d_out, d_img, 0, 0, d_edges);
}

// per-kernel sums of the last timed call, five slots: [0] band-energy spectra kernel (or the fused trace + energy
// kernel), [1] energy edges, [2] gain-application edges, [3] gain application, [4] trace pass when it ran on its own
int thz_chain_kernel_ms(const thz_ctx* c, float* ms5) {
  if (!c || !ms5) return THZ_EINVAL;
  for (int i = 0; i < 5; ++i) ms5[i] = c->kernel_ms[i];
  return THZ_OK;
}

// CUDA-event pairs around the cube kernels of ANY call between begin and end (the sharded path calls the passes
// one by one); end synchronises the context's streams and returns the sums like thz_deconv_kernel_ms.
int thz_kernel_timing_begin(thz_ctx* c) {
  CHECK_CTX(c);
  resolve_kernel_events(c);
  for (float& v : c->kernel_ms) v = 0.f;
  c->time_kernels = true;
  return THZ_OK;
}

int thz_kernel_timing_end(thz_ctx* c, float* ms4) {
  CHECK_CTX(c);
  c->time_kernels = false;
  THZ_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < kHostStreams; ++i) THZ_CUDA(c, cudaStreamSynchronize(c->hstream[i]));
  resolve_kernel_events(c);
  if (ms4)
    for (int i = 0; i < 4; ++i) ms4[i] = c->kernel_ms[i];
  return THZ_OK;
}

// The two cube-touching halves of thz_chain_host as separate calls, for hosts that run their own middle part
// (one process per GPU: thz_slab_rl between them).  begin leaves the filtered slab, its band energies
// [n_bands][P] and room for the gains [n_bands][P] on the device and returns the two pointers; end applies the
// gains found there and downloads.
int thz_chain_host_begin(thz_ctx* c, const float* cube, int rows, int cols, int n, const thz_band_plan* bands,
                         int n_bands, float* out, float** d_energy, float** d_gain) {
  CHECK_CTX(c);
  const int64_t P = (int64_t)rows * cols;
  if (P == 0) return THZ_OK;
  return chain_pass_in(c, cube, P, n, bands, n_bands, out, d_energy, d_gain);
}

int thz_chain_host_end(thz_ctx* c, int rows, int cols, int n, const thz_band_plan* bands, int n_bands, float* out,
                       float* img) {
  CHECK_CTX(c);
  const int64_t P = (int64_t)rows * cols;
  if (P == 0) return THZ_OK;
  if (!out) return set_err(c, THZ_EINVAL, "null pointer");
  return chain_pass_out(c, P, n, bands, n_bands, out, img);
}

int thz_deconv_stage_ms(const thz_ctx* c, float* ms4) {
  if (!c || !ms4) return THZ_EINVAL;
  for (int i = 0; i < 4; ++i) ms4[i] = c->stage_ms[i];
  return THZ_OK;
}

int thz_deconv_kernel_ms(const thz_ctx* c, float* ms4) {
  if (!c || !ms4) return THZ_EINVAL;
  for (int i = 0; i < 4; ++i) ms4[i] = c->kernel_ms[i];
  return THZ_OK;
}

int thz_deconvolution_host(thz_ctx* c, const float* cube, int rows, int cols, int n, const thz_band_plan* bands,
                           int n_bands, float* out, float* img, const volatile uint8_t* abort_flag,
                           thz_progress_fn progress, void* progress_user) {
  CHECK_CTX(c);
  const int64_t P = (int64_t)rows * cols;
  if (P == 0) return THZ_OK;
  if (!cube || !out) return set_err(c, THZ_EINVAL, "null pointer");
  void *pc = nullptr, *pi = nullptr;
  int rcw = ws_get(c, WS_HOST_CUBE, (size_t)P * n * sizeof(float), &pc);
  if (rcw == THZ_OK) rcw = ws_get(c, WS_HOST_IMG, (size_t)P * sizeof(float), &pi);
  if (rcw != THZ_OK) return rcw;
  float *d_cube = (float*)pc, *d_img = (float*)pi;
  THZ_CUDA(c, cudaMemcpyAsync(d_cube, cube, (size_t)P * n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  int rc = thz_deconvolution_dev(c, d_cube, rows, cols, n, bands, n_bands, d_cube, d_img, abort_flag, progress,
                                 progress_user);
  if (rc == THZ_OK) {
    cudaMemcpyAsync(out, d_cube, (size_t)P * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (img) cudaMemcpyAsync(img, d_img, (size_t)P * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
  }
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (rc == THZ_OK && e != cudaSuccess) return cuda_fail(c, e, "thz_deconvolution_host");
  return rc;
}

// The whole default chain + deconvolution from host memory to host memory with the cube resident on
// the device in between: H2D chunks overlap the fused trace pass and the band-energy pass, the gain
// application overlaps the D2H chunks; Richardson-Lucy runs in the middle on the B small images.
int thz_chain_host(thz_ctx* c, const float* cube, int rows, int cols, int n, const thz_band_plan* bands, int n_bands,
                   float* out, float* img, const volatile uint8_t* abort_flag, thz_progress_fn progress,
                   void* progress_user) {
  CHECK_CTX(c);
  const int64_t P = (int64_t)rows * cols;
  if (P == 0) return THZ_OK;
  if (progress) progress(0.0f, progress_user);
  float *d_energy = nullptr, *d_gain = nullptr;
  int rc = chain_pass_in(c, cube, P, n, bands, n_bands, out, &d_energy, &d_gain);
  if (rc != THZ_OK) return rc;
  if (n_bands) {
    rc = rl_all_bands(c, d_energy, P, rows, cols, bands, n_bands, d_gain, abort_flag, progress, progress_user);
    if (rc != THZ_OK) return rc;
  }
  rc = chain_pass_out(c, P, n, bands, n_bands, out, img);
  if (rc == THZ_OK && progress) progress(1.0f, progress_user);
  return rc;
}

}  // extern "C"
