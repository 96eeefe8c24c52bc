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
#include "thz_fft.cuh"
#include "thz_internal.h"

#include <math.h>
#include <algorithm>

namespace thz {

template <int M> struct DGeo {
  static constexpr int T = M / kE;
  static constexpr int NT = (T >= 256) ? T : 256;
  static constexpr int G = NT / T;
  static constexpr int kScr = (32 + kNzWords) * G;
  static constexpr size_t smem_bytes = (size_t)G * padded_len(M) * sizeof(float2) + kScr * sizeof(float);
  static constexpr int kMinBlocks = (NT == 256) ? 2 : 1;
};

struct FirArgs {
  const float* x;        // [P][N]
  int n;                 // samples per trace (<= M - 249)
  int64_t P;
  const float* hq;       // [B][M] zero-phase FIR spectra / M in last-stage register order
  int B;
  int64_t bstride;       // distance between bands in `energy` / `gain` (>= P; lets a chunk address a larger image)
  float* energy;         // [B][bstride]           (pass A)
  const float* gain;     // [B][bstride]           (pass C)
  float* out;            // [P][N]                 (pass C)
  float* img;            // [P] or null            (pass C)
  const float2* tw;
  const float* wq;       // [B][M/2] |H_b|^2 / M for the lower-half bins, register order   (pass A, Parseval)
  const float* wnyq;     // [B]      |H_b[M/2]|^2 / M
  const float2* edge;    // [2][B][512] spectra / 512 of the first / last 249 taps, register order (edges)
  const float2* tw512;
  // split form (M = 2N as two N-point sub-spectra, even and odd bins), all in Plan<N> register order
  const float* he;       // [B][N]   H_b[2j]   / M
  const float* ho;       // [B][N]   H_b[2j+1] / M
  const float* we;       // [B][N/2] |H_b[2j]|^2   / M, lower-half registers
  const float* wo;       // [B][N/2] |H_b[2j+1]|^2 / M
  const float2* mod;     // [N] exp(-2 pi i n / M): modulation that selects the odd bins
  float2* corr;          // [pairs][2][256] wrap-around corrections of the circular form (pass C)
  // band-interleaved copies of the lower-half tables (one 128-bit load serves four bands), Bp = B rounded up to 4
  int Bp;
  const float* we4;      // [Bp/4][N/2][4]
  const float* wo4;      // [Bp/4][N/2][4]
  const float* he4;      // [Bp/4][N/2][4]  H_b[2j] / M, lower-half registers
  const float* hny4;     // [Bp]       H_b[N] / M (Nyquist of the N-point spectrum)
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


// ------------------------------------------------------------------------------------
// Split form of the zero-padded M = 2N transform.  The trace occupies [0, N) of the M-point
// frame, so   X[2j]   = FFT_N(x)[j]               (even bins)
//             X[2j+1] = FFT_N(x * w_M^n)[j]       (odd bins, w_M = exp(-2 pi i / M))
// and for the outputs n < N:  y[n] = (inv_N(Y_even)[n] + conj(w_M^n) inv_N(Y_odd)[n]) / M.
// Mirror bins stay inside each sub-spectrum: M - 2j = 2 ((N - j) mod N), M - (2j+1) = 2 (N-1-j) + 1.
// Everything therefore runs in the 256-thread N-point geometry of the trace pass (2 CTAs per SM)
// instead of one 512-thread CTA per SM for a monolithic 8192-point transform.
// ------------------------------------------------------------------------------------
template <int N> struct SGeo {
  static constexpr int T = N / kE;
  static constexpr int NT = (T >= 256) ? T : 256;
  static constexpr int G = NT / T;
  static constexpr int kScr = (32 + kNzWords) * G;
  static constexpr size_t base_bytes = (size_t)G * padded_len(N) * sizeof(float2) + kScr * sizeof(float);
  // input slab of the work item, double buffered, filled by bulk copies (BulkStager): the even and the
  // odd pass both read it from shared memory, the cube crosses HBM once per pass
  static constexpr int kSlabFloats = 2 * G * N;
  static constexpr size_t stage_off = (base_bytes + 127) & ~(size_t)127;
  static constexpr size_t smem_bytes = stage_off + 2 * (size_t)kSlabFloats * sizeof(float) + 16;
  static constexpr int kMinBlocks = (NT == 256) ? 2 : 1;
};

template <int N>
__device__ __forceinline__ void load_pair_n(float2 (&v)[kE], const float* slab, int t, int g, bool act0, bool act1,
                                            bool& nz0, bool& nz1) {
  constexpr int T = SGeo<N>::T;
  const float* r0 = slab + (size_t)(2 * g) * N + t;
  const float* r1 = r0 + N;
  nz0 = nz1 = false;
#pragma unroll
  for (int i = 0; i < kE; ++i) {
    v[i].x = act0 ? r0[i * T] : 0.f;
    v[i].y = act1 ? r1[i * T] : 0.f;
    nz0 |= (v[i].x != 0.f);
    nz1 |= (v[i].y != 0.f);
  }
}

// unstaged variant: straight from global memory (streaming loads), leaves the shared memory to L1
template <int N>
__device__ __forceinline__ void load_pair_direct(float2 (&v)[kE], const float* __restrict__ x, int64_t p0, int t,
                                                 bool act0, bool act1, bool& nz0, bool& nz1) {
  constexpr int T = SGeo<N>::T;
  const float* r0 = x + p0 * N + t;
  const float* r1 = r0 + N;
  nz0 = nz1 = false;
#pragma unroll
  for (int i = 0; i < kE; ++i) {
    v[i].x = act0 ? __ldcs(r0 + i * T) : 0.f;
    v[i].y = act1 ? __ldcs(r1 + i * T) : 0.f;
    nz0 |= (v[i].x != 0.f);
    nz1 |= (v[i].y != 0.f);
  }
}

// common prologue / per-iteration staging of the split kernels
template <int N> struct SlabPipe {
  using GEO = SGeo<N>;
  float* slab[2];
  BulkStager stager;
  uint32_t it_count = 0;
  const float* x;
  int64_t P, nitems;
  __device__ __forceinline__ uint32_t bytes_of(int64_t it) const {
    int64_t cnt = P - it * GEO::G * 2;
    if (cnt > 2 * GEO::G) cnt = 2 * GEO::G;
    return (uint32_t)(cnt * N * sizeof(float));
  }
  __device__ __forceinline__ void init(unsigned char* smem_raw, const float* x_, int64_t P_, int64_t nitems_) {
    x = x_; P = P_; nitems = nitems_;
    slab[0] = reinterpret_cast<float*>(smem_raw + GEO::stage_off);
    slab[1] = slab[0] + GEO::kSlabFloats;
    stager.init(reinterpret_cast<uint64_t*>(smem_raw + GEO::stage_off + 2 * (size_t)GEO::kSlabFloats * sizeof(float)));
    if (threadIdx.x == 0 && (int64_t)blockIdx.x < nitems)
      stager.issue(0, slab[0], x + (int64_t)blockIdx.x * GEO::G * 2 * N, bytes_of(blockIdx.x));
  }
  // top of an iteration: prefetch the next item, wait for this one; returns its slab
  __device__ __forceinline__ const float* acquire(int64_t item) {
    const int buf = it_count & 1;
    const int64_t next = item + gridDim.x;
    if constexpr (GEO::T <= 32) __syncthreads();
    if (threadIdx.x == 0 && next < nitems)
      stager.issue(buf ^ 1, slab[buf ^ 1], x + next * GEO::G * 2 * N, bytes_of(next));
    stager.wait(buf, (it_count >> 1) & 1);
    ++it_count;
    return slab[buf];
  }
};

// x[n] * w_M^n for n = t + i*T: w_M^(t + i*T) = w_M^t * exp(-i pi i / 16) since T / M = 1 / 32, so one table entry
// per thread (2 KB of the table stay hot instead of 32 KB) and 16 compile-time rotations
template <int N>
__device__ __forceinline__ void modulate(float2 (&v)[kE], const float2* __restrict__ mod, int t) {
  static_assert(kE == 16, "rotation constants are exp(-i pi i / 16)");
  constexpr float kC[16] = {1.0f, 0.98078528f, 0.923879533f, 0.831469612f, 0.707106781f, 0.555570233f, 0.382683432f, 0.195090322f, 0.0f, -0.195090322f, -0.382683432f, -0.555570233f, -0.707106781f, -0.831469612f, -0.923879533f, -0.98078528f};
  constexpr float kS[16] = {0.0f, -0.195090322f, -0.382683432f, -0.555570233f, -0.707106781f, -0.831469612f, -0.923879533f, -0.98078528f, -1.0f, -0.98078528f, -0.923879533f, -0.831469612f, -0.707106781f, -0.555570233f, -0.382683432f, -0.195090322f};
  const float2 w0 = __ldg(mod + t);
#pragma unroll
  for (int i = 0; i < kE; ++i) {
    const float2 w = make_float2(w0.x * kC[i] - w0.y * kS[i], w0.x * kS[i] + w0.y * kC[i]);
    v[i] = cmul(v[i], w);
  }
}

// q1 = p + c, q2 = p - c for the lower-half registers of one sub-spectrum (see k_fir_energy_total)
template <int N, bool ODD>
__device__ __forceinline__ void parseval_terms(const float2 (&z)[kE], const float2* sm, int t, float (&q1)[kE / 2],
                                               float (&q2)[kE / 2]) {
  constexpr int LAST = Plan<N>::ns - 1;
#pragma unroll
  for (int j = 0; j < kE / 2; ++j) {
    const int f = pos_to_bin<N>(stage_elem<N, LAST>(t, j));
    const float2 zz = z[j];
    if (!ODD && f == 0) {
      q1[j] = zz.x * zz.x;
      q2[j] = zz.y * zz.y;
    } else {
      const float2 zp = sm[pad_idx(ODD ? (N - 1 - f) : (N - f))];
      const float p = 0.5f * (zz.x * zz.x + zz.y * zz.y + zp.x * zp.x + zp.y * zp.y);
      const float c = zz.x * zp.x - zz.y * zp.y;
      q1[j] = p + c;
      q2[j] = p - c;
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
  constexpr int W = (T < 32) ? T : 32;
  constexpr int NW = (T + 31) / 32;
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
    // eight bands per round: e[2*bb + trace]; the 16 partial sums of a warp are reduced by a halving tree
    // (16 shuffles instead of 80), then across the warps through shared memory
    for (int b0 = 0; b0 < a.B; b0 += 8) {
      float e[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) e[k] = 0.f;
#pragma unroll
      for (int j = 0; j < NLOW; ++j) {
        const int u = j % UL, m = j / UL;
        const int idx = m * (N / RL) + t + u * T;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (b0 + 4 * h < a.Bp) {
            const size_t o = ((size_t)((b0 >> 2) + h) * (N / 2) + idx) * 4;
            const float4 w_e = __ldg(reinterpret_cast<const float4*>(a.we4 + o));
            const float4 w_o = __ldg(reinterpret_cast<const float4*>(a.wo4 + o));
            const float we_[4] = {w_e.x, w_e.y, w_e.z, w_e.w}, wo_[4] = {w_o.x, w_o.y, w_o.z, w_o.w};
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
              const int k = 2 * (4 * h + bb);
              e[k] = fmaf(we_[bb], q1e[j], e[k]);
              e[k + 1] = fmaf(we_[bb], q2e[j], e[k + 1]);
              e[k] = fmaf(wo_[bb], q1o[j], e[k]);
              e[k + 1] = fmaf(wo_[bb], q2o[j], e[k + 1]);
            }
          }
        }
      }
      if (t == 0) {
#pragma unroll
        for (int bb = 0; bb < 8; ++bb) {
          if (b0 + bb < a.B) {
            const float w = __ldg(a.wnyq + b0 + bb);
            e[2 * bb] = fmaf(w, ny1, e[2 * bb]);
            e[2 * bb + 1] = fmaf(w, ny2, e[2 * bb + 1]);
          }
        }
      }
      if constexpr (W == 32) {
        // after the step with lane offset `off` a lane keeps the upper half of its values when its `off` bit is
        // set: lane l ends with the warp total of value l >> 1
        const int lane = t & 31;
#pragma unroll
        for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
          const bool up = (lane & off) != 0;
#pragma unroll
          for (int k = 0; k < half; ++k) {
            const float send = up ? e[k] : e[k + half];
            const float keep = up ? e[k + half] : e[k];
            e[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        e[0] += __shfl_xor_sync(0xffffffffu, e[0], 1);
        if constexpr (T > 32) {
          if ((lane & 1) == 0) red[(t >> 5) * 16 + (lane >> 1)] = e[0];
          __syncthreads();
          if (t < 16) {   // thread t sums value t = (band t/2, trace t%2) over the warps
            float acc = 0.f;
            for (int w = 0; w < NW; ++w) acc += red[w * 16 + t];
            const int bnd = b0 + (t >> 1);
            const bool second = (t & 1) != 0;
            if (bnd < a.B && (second ? act1 : act0))
              a.energy[(size_t)bnd * a.bstride + p0 + (second ? 1 : 0)] = (second ? z1 : z0) ? 0.f : acc;
          }
          __syncthreads();
        } else {
          // one warp per pair: lane l holds value l >> 1
          const int v = lane >> 1, bnd = b0 + (v >> 1);
          const bool second = (v & 1) != 0;
          if ((lane & 1) == 0 && bnd < a.B && (second ? act1 : act0))
            a.energy[(size_t)bnd * a.bstride + p0 + (second ? 1 : 0)] = (second ? z1 : z0) ? 0.f : e[0];
        }
      } else {
        // groups narrower than a warp (N < 512): plain butterfly inside the group
#pragma unroll
        for (int k = 0; k < 16; ++k) {
#pragma unroll
          for (int o = W / 2; o > 0; o >>= 1) e[k] += __shfl_xor_sync(0xffffffffu, e[k], o);
        }
        if (t == 0) {
#pragma unroll
          for (int bb = 0; bb < 8; ++bb) {
            const int bnd = b0 + bb;
            if (bnd < a.B) {
              if (act0) a.energy[(size_t)bnd * a.bstride + p0] = z0 ? 0.f : e[2 * bb];
              if (act1) a.energy[(size_t)bnd * a.bstride + p0 + 1] = z1 ? 0.f : e[2 * bb + 1];
            }
          }
        }
      }
    }
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
template <int N>
__device__ __forceinline__ void mix_paired(float2 (&z)[kE], float2* sm, int t, const FirArgs& a,
                                           const float* __restrict__ htab, int64_t p0, bool act0, bool act1,
                                           bool& bad0, bool& bad1, float gmul) {
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
      float g0 = (act0 && b < a.B) ? __ldg(a.gain + (size_t)b * a.bstride + p0) : 0.f;
      float g1 = (act1 && b < a.B) ? __ldg(a.gain + (size_t)b * a.bstride + p0 + 1) : 0.f;
      if (!(fabsf(g0) <= 3.0e38f)) { bad0 = true; g0 = 0.f; }
      if (!(fabsf(g1) <= 3.0e38f)) { bad1 = true; g1 = 0.f; }
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
        for (int b = 0; b < a.B; ++b) {
          float g0 = __ldg(a.gain + (size_t)b * a.bstride + p0);
          float g1 = act1 ? __ldg(a.gain + (size_t)b * a.bstride + p0 + 1) : 0.f;
          if (!(fabsf(g0) <= 3.0e38f)) g0 = 0.f;   // the main kernel marks such traces NaN
          if (!(fabsf(g1) <= 3.0e38f)) g1 = 0.f;
          const float gs = 0.5f * (g0 + g1), gd = 0.5f * (g0 - g1);
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

template <int N, bool STAGED>
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
    if constexpr (STAGED) {
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
    fft_forward<N>(z, t, sm, a.tw);
    __syncthreads();
#pragma unroll
    for (int i = kE / 2; i < kE; ++i) sm[pad_idx(pos_to_bin<N>(stage_elem<N, LAST>(t, i)))] = z[i];   // mirrors of lower-half bins are upper-half registers
    __syncthreads();
    // he holds H / (2N): the N-point inverse needs H / N
    mix_paired<N>(z, sm, t, a, a.he, p0, act0, act1, bad0, bad1, 1.0f);
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
// Richardson-Lucy: TMA-staged tiled 2-D filtering
// ------------------------------------------------------------------------------------
constexpr int kTH = 64, kTW = 64;          // output tile
constexpr int kMidStride = 68;             // 4 * odd -> conflict-free 128-bit rows
constexpr int kMaxTaps = 256;              // padded taps per axis

struct ConvArgs {
  int Hp, Wp, pitch;      // padded-domain image [Hp][pitch], valid width Wp
  int kx, ky;             // taps along rows (axis 0) and columns (axis 1), both odd
  int kxp, kyp;           // taps padded to a multiple of 8
  int box_rows, box_cols; // TMA box: kTH + kx - 1 rows, >= kTW + kyp - 1 columns (4 * odd)
  const float* wx;        // [kxp] row-direction taps (correlation order), zero padded
  const float* wy;        // [kyp]
  const float* wdense;    // [kx][kyp] dense taps (dense kernel) or null
  const float* d;         // mode 1: relative blur numerator (padded image)
  float* out;             // mode 0: conv result; mode 1: r = d / (conv + eps); mode 2: u *= conv (in place)
  float eps;
  int col_shift;          // tile-grid column offset that keeps the TMA box start 16-byte aligned
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t mbar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(mbar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t mbar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(mbar)
      : "memory");
}

// correlation: out[i][j] = sum_m sum_n in[i + m - kx/2][j + n - ky/2] * wx[m] * wy[n]
// MODE 0: out = c;  MODE 1: out = d / (c + eps);  MODE 2: out *= c
template <int MODE, bool DENSE>
__global__ void __launch_bounds__(256, 2) k_rl_conv(const __grid_constant__ CUtensorMap tmap, const ConvArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // TMA destinations must be 128-byte aligned: align by hand (the launch adds 128 bytes of slack),
  // the mbarrier lives in the first 16 bytes of the aligned block
  unsigned char* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint64_t* mbar_ptr = reinterpret_cast<uint64_t*>(base);
  float* tile = reinterpret_cast<float*>(base + 128);                     // [box_rows][box_cols]
  const int tile_floats = a.box_rows * a.box_cols;
  float* mid = tile + ((tile_floats + 31) & ~31);                         // [box_rows][kMidStride]
  float* wxs = mid + (DENSE ? 0 : a.box_rows * kMidStride);
  float* wys = wxs + (DENSE ? 0 : a.kxp);                                 // separable: [kxp] then [kyp]
  const uint32_t mbar = smem_u32(mbar_ptr);
  // TMA needs a 16-byte aligned box start: the innermost coordinate col0 - ky/2 must be a multiple of
  // 4 floats, so the tile grid is shifted left by col_shift in {0, -3, -2, -1} columns
  const int row0 = blockIdx.y * kTH, col0 = blockIdx.x * kTW + a.col_shift;

  if (threadIdx.x == 0) mbar_init(mbar, 1);
  if constexpr (DENSE) {
    for (int i = threadIdx.x; i < a.kx * a.kyp; i += blockDim.x) wxs[i] = a.wdense[i];
  } else {
    for (int i = threadIdx.x; i < a.kxp; i += blockDim.x) wxs[i] = a.wx[i];
    for (int i = threadIdx.x; i < a.kyp; i += blockDim.x) wys[i] = a.wy[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(mbar, (uint32_t)(tile_floats * sizeof(float)));
    tma_load_2d(smem_u32(tile), &tmap, col0 - a.ky / 2, row0 - a.kx / 2, mbar);
  }
  while (!mbar_try_wait(mbar, 0)) {
  }

  const int bc = a.box_cols;
  if constexpr (!DENSE) {
    // pass 1: filter along columns (axis 1).  item = (tile row r, group of 8 output columns)
    const int nitems = a.box_rows * (kTW / 8);
    for (int it = threadIdx.x; it < nitems; it += blockDim.x) {
      const int r = it % a.box_rows, cg = it / a.box_rows;
      const float* src = tile + r * bc + cg * 8;
      float acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
      float win[16];
      {
        const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
        win[0] = v0.x; win[1] = v0.y; win[2] = v0.z; win[3] = v0.w;
        win[4] = v1.x; win[5] = v1.y; win[6] = v1.z; win[7] = v1.w;
      }
      for (int nb = 0; nb < a.kyp; nb += 8) {
        const float4 v0 = *reinterpret_cast<const float4*>(src + nb + 8);
        const float4 v1 = *reinterpret_cast<const float4*>(src + nb + 12);
        win[8] = v0.x; win[9] = v0.y; win[10] = v0.z; win[11] = v0.w;
        win[12] = v1.x; win[13] = v1.y; win[14] = v1.z; win[15] = v1.w;
        const float4 w0 = *reinterpret_cast<const float4*>(wys + nb), w1 = *reinterpret_cast<const float4*>(wys + nb + 4);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[q] = fmaf(wv[n], win[q + n], acc[q]);
#pragma unroll
        for (int q = 0; q < 8; ++q) win[q] = win[q + 8];
      }
      float* dst = mid + r * kMidStride + cg * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    __syncthreads();
  }
  // pass 2: filter along rows (axis 0).  item = (output column c, group of 8 output rows)
  {
    const int nitems = kTW * (kTH / 8);
    for (int it = threadIdx.x; it < nitems; it += blockDim.x) {
      const int c = it % kTW, rg = it / kTW;
      float acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
      if constexpr (!DENSE) {
        const float* src = mid + (rg * 8) * kMidStride + c;
        float win[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) win[q] = src[q * kMidStride];
        for (int mb = 0; mb < a.kxp; mb += 8) {
          // rows beyond the box are multiplied by zero taps; clamp the address instead of reading them
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int rr = rg * 8 + mb + 8 + q;
            win[8 + q] = (rr < a.box_rows) ? mid[rr * kMidStride + c] : 0.f;
          }
          const float4 w0 = *reinterpret_cast<const float4*>(wxs + mb), w1 = *reinterpret_cast<const float4*>(wxs + mb + 4);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = fmaf(wv[n], win[q + n], acc[q]);
#pragma unroll
          for (int q = 0; q < 8; ++q) win[q] = win[q + 8];
        }
      } else {
        // dense taps: out[r][c] = sum_m sum_n tile[r + m][c + n] w[m][n]; 8 consecutive rows per item,
        // tap rows outermost so that one tap value serves 8 accumulators
        for (int m = 0; m < a.kx + 7; ++m) {
          // input row rg*8 + m contributes to output row q with tap row m - q
          const float* srow = tile + (rg * 8 + m) * bc + c;
          if (rg * 8 + m >= a.box_rows) break;
          for (int n = 0; n < a.ky; ++n) {
            const float v = srow[n];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int mq = m - q;
              if (mq >= 0 && mq < a.kx) acc[q] = fmaf(v, wxs[mq * a.kyp + n], acc[q]);
            }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int gr = row0 + rg * 8 + q, gc = col0 + c;
        if (gr < a.Hp && gc >= 0 && gc < a.Wp) {
          const size_t o = (size_t)gr * a.pitch + gc;
          if constexpr (MODE == 0) a.out[o] = acc[q];
          else if constexpr (MODE == 1) a.out[o] = a.d[o] / (acc[q] + a.eps);
          else a.out[o] = a.out[o] * acc[q];
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------
// Persistent separable RL filtering: one CTA per SM loops over the 64x64 output tiles; the haloed
// input tile of the NEXT tile is fetched by TMA into the other buffer while the current one is
// filtered (column pass -> row pass -> fused epilogue).  512 threads: 880 column-pass items
// (row, 8 columns), 512 row-pass items (column, 8 rows).
// ------------------------------------------------------------------------------------
constexpr int kRlThreads = 512;

template <int MODE>
__global__ void __launch_bounds__(kRlThreads, 1) k_rl_conv_persistent(const __grid_constant__ CUtensorMap tmap,
                                                                      const ConvArgs a, int tiles_x, int tiles_y) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  uint64_t* mbar_ptr = reinterpret_cast<uint64_t*>(base);   // two barriers
  const int tile_floats = (a.box_rows * a.box_cols + 31) & ~31;
  float* tile0 = reinterpret_cast<float*>(base + 128);
  float* tile1 = tile0 + tile_floats;
  float* mid = tile1 + tile_floats;                           // [box_rows][kMidStride]
  float* wxs = mid + a.box_rows * kMidStride;
  float* wys = wxs + a.kxp;
  const uint32_t mbar[2] = {smem_u32(mbar_ptr), smem_u32(mbar_ptr + 1)};
  const int ntiles = tiles_x * tiles_y;
  const uint32_t tile_bytes = (uint32_t)(a.box_rows * a.box_cols * sizeof(float));

  if (threadIdx.x == 0) {
    mbar_init(mbar[0], 1);
    mbar_init(mbar[1], 1);
  }
  for (int i = threadIdx.x; i < a.kxp; i += blockDim.x) wxs[i] = a.wx[i];
  for (int i = threadIdx.x; i < a.kyp; i += blockDim.x) wys[i] = a.wy[i];
  __syncthreads();
  auto origin = [&](int tidx, int& row0, int& col0) {
    row0 = (tidx / tiles_x) * kTH;
    col0 = (tidx % tiles_x) * kTW + a.col_shift;
  };
  if (threadIdx.x == 0 && (int)blockIdx.x < ntiles) {
    int r0, c0;
    origin(blockIdx.x, r0, c0);
    mbar_expect_tx(mbar[0], tile_bytes);
    tma_load_2d(smem_u32(tile0), &tmap, c0 - a.ky / 2, r0 - a.kx / 2, mbar[0]);
  }
  const int bc = a.box_cols;
  uint32_t it = 0;
  for (int tidx = blockIdx.x; tidx < ntiles; tidx += gridDim.x, ++it) {
    const int buf = it & 1;
    float* tile = buf ? tile1 : tile0;
    int row0, col0;
    origin(tidx, row0, col0);
    const int nxt = tidx + gridDim.x;
    if (threadIdx.x == 0 && nxt < ntiles) {   // the other buffer was last read before the barrier that ended
      int r0, c0;                              // the previous iteration
      origin(nxt, r0, c0);
      mbar_expect_tx(mbar[buf ^ 1], tile_bytes);
      tma_load_2d(smem_u32(buf ? tile0 : tile1), &tmap, c0 - a.ky / 2, r0 - a.kx / 2, mbar[buf ^ 1]);
    }
    // epilogue operands: fetch early so that their latency hides behind the column pass
    const int c = threadIdx.x % kTW, rg = threadIdx.x / kTW;
    float ep[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int gr = row0 + rg * 8 + q, gc = col0 + c;
      ep[q] = 0.f;
      if (MODE != 0 && gr < a.Hp && gc >= 0 && gc < a.Wp) {
        const size_t o = (size_t)gr * a.pitch + gc;
        ep[q] = (MODE == 1) ? __ldg(a.d + o) : a.out[o];
      }
    }
    while (!mbar_try_wait(mbar[buf], (it >> 1) & 1)) {
    }
    // column pass (axis 1)
    const int nitems = a.box_rows * (kTW / 8);
    for (int itx = threadIdx.x; itx < nitems; itx += blockDim.x) {
      const int r = itx % a.box_rows, cg = itx / a.box_rows;
      const float* src = tile + r * bc + cg * 8;
      float acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
      float win[16];
      {
        const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 4);
        win[0] = v0.x; win[1] = v0.y; win[2] = v0.z; win[3] = v0.w;
        win[4] = v1.x; win[5] = v1.y; win[6] = v1.z; win[7] = v1.w;
      }
      for (int nb = 0; nb < a.kyp; nb += 8) {
        const float4 v0 = *reinterpret_cast<const float4*>(src + nb + 8);
        const float4 v1 = *reinterpret_cast<const float4*>(src + nb + 12);
        win[8] = v0.x; win[9] = v0.y; win[10] = v0.z; win[11] = v0.w;
        win[12] = v1.x; win[13] = v1.y; win[14] = v1.z; win[15] = v1.w;
        const float4 w0 = *reinterpret_cast<const float4*>(wys + nb), w1 = *reinterpret_cast<const float4*>(wys + nb + 4);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[q] = fmaf(wv[n], win[q + n], acc[q]);
#pragma unroll
        for (int q = 0; q < 8; ++q) win[q] = win[q + 8];
      }
      float* dst = mid + r * kMidStride + cg * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    __syncthreads();
    // row pass (axis 0): one item per thread
    {
      float acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
      const float* src = mid + (rg * 8) * kMidStride + c;
      float win[16];
#pragma unroll
      for (int q = 0; q < 8; ++q) win[q] = src[q * kMidStride];
      for (int mb = 0; mb < a.kxp; mb += 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int rr = rg * 8 + mb + 8 + q;
          win[8 + q] = (rr < a.box_rows) ? mid[rr * kMidStride + c] : 0.f;
        }
        const float4 w0 = *reinterpret_cast<const float4*>(wxs + mb), w1 = *reinterpret_cast<const float4*>(wxs + mb + 4);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[q] = fmaf(wv[n], win[q + n], acc[q]);
#pragma unroll
        for (int q = 0; q < 8; ++q) win[q] = win[q + 8];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int gr = row0 + rg * 8 + q, gc = col0 + c;
        if (gr < a.Hp && gc >= 0 && gc < a.Wp) {
          const size_t o = (size_t)gr * a.pitch + gc;
          if constexpr (MODE == 0) a.out[o] = acc[q];
          else if constexpr (MODE == 1) a.out[o] = ep[q] / (acc[q] + a.eps);
          else a.out[o] = ep[q] * acc[q];
        }
      }
    }
    __syncthreads();   // tile[buf] and mid are free again
  }
}

// ------------------------------------------------------------------------------------
// Streaming separable RL filtering.  A CTA owns a strip of kSW output columns and a segment of
// output rows and marches down the strip in chunks of kCR input rows: each chunk is fetched by TMA
// (two buffers, prefetch distance two chunks), filtered along the columns into a ring of
// column-filtered rows kept in shared memory (stored transposed, so that the row pass reads its
// taps' axis with 128-bit loads), and the output rows whose whole row support is in the ring are
// emitted with the fused epilogue.  Every image row is column-filtered once per segment (only the
// WU warm-up rows of a segment are filtered twice), both passes are exactly one item per thread
// (16 outputs x T taps), and the ring holds one chunk of slack so that one barrier per chunk suffices.
//
// Taps are front-padded with zeros to WU + 1 (rows) / KW + 1 (columns) entries, WU and KW being the
// true support minus one rounded up to a multiple of 8, so that the window advances in whole
// 8-element blocks and the last tap is a single trailing step that needs no new data.
// ------------------------------------------------------------------------------------
constexpr int kCR = 64;           // input rows per chunk
// SW = output columns per strip (128: one 512-thread CTA per SM; 64: 256 threads, two CTAs per SM when the
// buffers fit twice, so that one CTA computes while the other sits at its barrier); SW * kCR / 16 threads:
// one 16-output item per thread in both passes

struct StreamArgs {
  int Hp, Wp, pitch;
  int WU, KW;           // warm-up rows / columns (multiples of 8)
  int gy_off, gx_off;   // box origin = (segment row 0 - gy_off, strip column 0 - gx_off)
  int bc;               // box columns (4 * odd, >= SW + KW)
  int Rg, RS;           // ring rows (multiple of 8, >= 2 kCR + WU) and ring stride in floats (4 * odd)
  int seg_rows;         // output rows per segment (kCR * chunks - WU)
  int col_shift;        // keeps the box start 16-byte aligned
  const float* wx;      // [WU + 8] front-padded row taps, wx[WU] is the last tap
  const float* wy;      // [KW + 8]
  const float* d;
  float* out;
  float eps;
};

// logical window element idx in [0, 24) -> register, the three 8-groups rotate with the phase
__device__ __forceinline__ constexpr int win_phys(int idx, int ph) { return (((idx >> 3) + ph) % 3) * 8 + (idx & 7); }

template <int PH>
__device__ __forceinline__ void tap_block(float (&acc)[16], float (&W)[24], const float4 n0, const float4 n1,
                                          const float* __restrict__ w8) {
  constexpr int g2 = ((2 + PH) % 3) * 8;
  W[g2 + 0] = n0.x; W[g2 + 1] = n0.y; W[g2 + 2] = n0.z; W[g2 + 3] = n0.w;
  W[g2 + 4] = n1.x; W[g2 + 5] = n1.y; W[g2 + 6] = n1.z; W[g2 + 7] = n1.w;
  const float4 wa = *reinterpret_cast<const float4*>(w8), wb = *reinterpret_cast<const float4*>(w8 + 4);
  const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[q] = fmaf(wv[n], W[win_phys(q + n, PH)], acc[q]);
}

template <int PH>
__device__ __forceinline__ void tap_last(float (&acc)[16], const float (&W)[24], float wl) {
#pragma unroll
  for (int q = 0; q < 16; ++q) acc[q] = fmaf(wl, W[win_phys(q, PH)], acc[q]);
}

// 16 outputs, T + 1 taps (T a multiple of 8).  L::next() returns the next 8 window elements.
template <class L>
__device__ __forceinline__ void run_taps(float (&acc)[16], L& ld, const float* __restrict__ w, int T) {
  float W[24];
  float4 a, b;
  ld.next(a, b);
  W[0] = a.x; W[1] = a.y; W[2] = a.z; W[3] = a.w; W[4] = b.x; W[5] = b.y; W[6] = b.z; W[7] = b.w;
  ld.next(a, b);
  W[8] = a.x; W[9] = a.y; W[10] = a.z; W[11] = a.w; W[12] = b.x; W[13] = b.y; W[14] = b.z; W[15] = b.w;
  int nb = 0;
  for (; nb + 24 <= T; nb += 24) {
    ld.next(a, b);
    tap_block<0>(acc, W, a, b, w + nb);
    ld.next(a, b);
    tap_block<1>(acc, W, a, b, w + nb + 8);
    ld.next(a, b);
    tap_block<2>(acc, W, a, b, w + nb + 16);
  }
  const int rem = (T - nb) >> 3;
  const float wl = w[T];
  if (rem == 0) {
    tap_last<0>(acc, W, wl);
  } else if (rem == 1) {
    ld.next(a, b);
    tap_block<0>(acc, W, a, b, w + nb);
    tap_last<1>(acc, W, wl);
  } else {
    ld.next(a, b);
    tap_block<0>(acc, W, a, b, w + nb);
    ld.next(a, b);
    tap_block<1>(acc, W, a, b, w + nb + 8);
    tap_last<2>(acc, W, wl);
  }
}

struct LinearLoader {   // consecutive floats of one haloed tile row
  const float* p;
  __device__ __forceinline__ void next(float4& a, float4& b) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
    p += 8;
  }
};

struct RingLoader {     // consecutive ring rows of one column (transposed ring: rows are contiguous)
  const float* col;
  int pos, Rg;
  __device__ __forceinline__ void next(float4& a, float4& b) {
    a = *reinterpret_cast<const float4*>(col + pos);
    b = *reinterpret_cast<const float4*>(col + pos + 4);
    pos += 8;
    if (pos >= Rg) pos -= Rg;
  }
};

template <int MODE, int SW>
__global__ void __launch_bounds__(SW * 4, (SW == 64) ? 2 : 1) k_rl_stream(const __grid_constant__ CUtensorMap tmap,
                                                                         const StreamArgs a) {
  constexpr int kSW = SW, kStThreads = SW * 4;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
  const uint32_t mbar0 = smem_u32(base);                       // two barriers, 8 bytes apart
  const int tile_floats = kCR * a.bc;
  float* tile0 = reinterpret_cast<float*>(base + 128);
  float* ring = tile0 + 2 * tile_floats;                       // [kSW][RS]
  float* wxs = ring + kSW * a.RS;
  float* wys = wxs + a.WU + 8;
  const int tid = threadIdx.x;
  const int col0 = blockIdx.x * kSW + a.col_shift;
  const int seg_row0 = blockIdx.y * a.seg_rows;
  const int rows_out = min(a.seg_rows, a.Hp - seg_row0);
  if (rows_out <= 0) return;
  const int n_chunks = (rows_out + a.WU + kCR - 1) / kCR;
  const int gy0 = seg_row0 - a.gy_off, gx0 = col0 - a.gx_off;
  const uint32_t tile_bytes = (uint32_t)(tile_floats * sizeof(float));
  // a chunk that lies wholly above or below the image is all zeros: no copy, the ring rows are cleared
  auto live = [&](int j) { return gy0 + kCR * j < a.Hp && gy0 + kCR * (j + 1) > 0; };

  // programmatic dependent launch: the next kernel of the iteration may be scheduled as soon as every CTA
  // of this one is running, its prologue (barriers, taps) overlaps our tail; everything that touches the
  // images comes after griddepcontrol.wait, which returns when the previous kernel has completed
  asm volatile("griddepcontrol.launch_dependents;");
  if (tid == 0) {
    mbar_init(mbar0, 1);
    mbar_init(mbar0 + 8, 1);
  }
  for (int i = tid; i < a.WU + 8; i += kStThreads) wxs[i] = a.wx[i];
  for (int i = tid; i < a.KW + 8; i += kStThreads) wys[i] = a.wy[i];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    for (int j = 0; j < 2 && j < n_chunks; ++j)
      if (live(j)) {
        mbar_expect_tx(mbar0 + 8 * j, tile_bytes);
        tma_load_2d(smem_u32(tile0 + j * tile_floats), &tmap, gx0, gy0 + kCR * j, mbar0 + 8 * j);
      }
  }
  uint32_t ph0 = 0, ph1 = 0;   // mbarrier phase parity per buffer
  for (int j = 0; j < n_chunks; ++j) {
    const int buf = j & 1;
    const bool lv = live(j);
    // rows that become complete with this chunk, and this thread's share of them in the row pass:
    // thread = (strip column c, 16 output rows).  The epilogue operands are fetched now, so that their
    // latency hides behind the column pass even when the PSF is small.
    const int lo = max(0, j * kCR - a.WU), hi = min(rows_out, (j + 1) * kCR - a.WU);
    const int c = tid & (kSW - 1), rg = tid / kSW;
    const int i0 = lo + rg * 16;
    const int gc = col0 + c;
    const bool colok = gc >= 0 && gc < a.Wp;
    const size_t o0 = (size_t)(seg_row0 + i0) * a.pitch + gc;
    float ep[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      ep[q] = 0.f;
      if (MODE != 0 && colok && i0 + q < hi) {
        const size_t o = o0 + (size_t)q * a.pitch;
        ep[q] = (MODE == 1) ? __ldg(a.d + o) : a.out[o];
      }
    }
    // ---- column pass: thread = (chunk row r, 16 output columns cg*16 ..) ----
    {
      const int r = tid & (kCR - 1), cg = tid >> 6;
      float acc[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) acc[q] = 0.f;
      if (lv) {
        const uint32_t mb = mbar0 + 8 * buf;
        while (!mbar_try_wait(mb, buf ? ph1 : ph0)) {
        }
        if (buf) ph1 ^= 1; else ph0 ^= 1;
        LinearLoader ld{tile0 + buf * tile_floats + r * a.bc + cg * 16};
        run_taps(acc, ld, wys, a.KW);
      }
      int pos = (j * kCR + r) % a.Rg;
      float* dst = ring + (cg * 16) * a.RS + pos;
#pragma unroll
      for (int q = 0; q < 16; ++q) dst[q * a.RS] = acc[q];
    }
    __syncthreads();   // ring rows of chunk j are visible; tile[buf] is free
    if (tid == 0 && j + 2 < n_chunks && live(j + 2)) {
      mbar_expect_tx(mbar0 + 8 * buf, tile_bytes);
      tma_load_2d(smem_u32(tile0 + buf * tile_floats), &tmap, gx0, gy0 + kCR * (j + 2), mbar0 + 8 * buf);
    }
    // ---- row pass ----
    if (i0 < hi) {
      float acc[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) acc[q] = 0.f;
      RingLoader ld{ring + c * a.RS, i0 % a.Rg, a.Rg};
      run_taps(acc, ld, wxs, a.WU);
      if (colok) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          if (i0 + q < hi) {
            const size_t o = o0 + (size_t)q * a.pitch;
            if constexpr (MODE == 0) a.out[o] = acc[q];
            else if constexpr (MODE == 1) a.out[o] = ep[q] / (acc[q] + a.eps);
            else a.out[o] = ep[q] * acc[q];
          }
        }
      }
    }
  }
}

// numpy-"reflect" padding exactly as richardson_lucy writes it (deconvolution.rs:638-667)
__global__ void k_reflect_pad(const float* __restrict__ img, int h, int w, int pad_y, int pad_x, float* __restrict__ out,
                              int pitch) {
  const int Hp = h + 2 * pad_y, Wp = w + 2 * pad_x;
  const int64_t total = (int64_t)Hp * Wp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / Wp), c = (int)(i % Wp);
    int sr, sc;
    if (r < pad_y) sr = pad_y - r;
    else if (r >= pad_y + h) sr = h - 2 - (r - pad_y - h);
    else sr = r - pad_y;
    if (c < pad_x) sc = pad_x - c;
    else if (c >= pad_x + w) sc = w - 2 - (c - pad_x - w);
    else sc = c - pad_x;
    out[(size_t)r * pitch + c] = img[(size_t)sr * w + sc];
  }
}

// crop, clamp >= 0, gain = sqrt(u / d)  (deconvolution.rs:708, 975, 990-993)
__global__ void k_rl_finish(const float* __restrict__ u, int pitch, int pad_y, int pad_x, int h, int w,
                            const float* __restrict__ d_img, float* __restrict__ deconv, float* __restrict__ gain) {
  const int64_t total = (int64_t)h * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / w), c = (int)(i % w);
    const float v = fmaxf(u[(size_t)(r + pad_y) * pitch + c + pad_x], 0.0f);
    if (deconv) deconv[i] = v;
    if (gain) gain[i] = sqrtf(v / d_img[i]);
  }
}

__global__ void k_copy2d(const float* __restrict__ src, int rows, int cols, int spitch, float* __restrict__ dst,
                         int dpitch) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    dst[(size_t)r * dpitch + c] = src[(size_t)r * spitch + c];
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
                    int B, float* d_energy, int64_t bstride = 0) {
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
        rc = launch_fir<512>(c, s, k_fir_edges, a);
        kt.stop();
      }
    } else {
      rc = dispatch_energy(c, s, ft.m, a);   // short traces: B inverse transforms per pair
    }
  }
  return rc;
}

// `lane` selects the wrap-around correction workspace: calls that are in flight on different streams at the same
// time (the chunk pipeline of thz_chain_host) must not share one (k_fir_edge_corr of chunk i+1 would overwrite
// what k_fir_apply_circ of chunk i still reads).  Lane 0 = the context's compute stream, 1 + k = hstream[k].
int deconv_apply(thz_ctx* c, cudaStream_t s, const float* d_cube, const float* d_gain, int64_t P, int n,
                 const thz_band_plan* bands, int B, float* d_out, float* d_img, int64_t bstride = 0, int lane = 0) {
  if (bstride == 0) bstride = P;
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
            KernelTimer kt(c, s, 2);
            rc = launch_fir<512>(c, s, k_fir_edge_corr, ac);
            kt.stop();
          }
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

// ---- TMA descriptor -----------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode(thz_ctx* c) {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    set_err(c, THZ_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    return nullptr;
  }
  fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  return fn;
}

static int make_tmap(thz_ctx* c, CUtensorMap* map, const float* base, int Hp, int Wp, int pitch, int box_rows,
                     int box_cols) {
  auto enc = get_encode(c);
  if (!enc) return THZ_ECUDA;
  cuuint64_t dims[2] = {(cuuint64_t)Wp, (cuuint64_t)Hp};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[128];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return set_err(c, THZ_ECUDA, buf);
  }
  return THZ_OK;
}

struct ConvPlan {
  ConvArgs a{};
  size_t smem = 0;
  bool dense = false;
  float* d_w = nullptr;   // device taps: [wx (kxp) | wy (kyp)] x 2 orientations, or dense [2][kx][kyp]
  int wstride = 0;        // floats between the two orientations
  // streaming form (k_rl_stream): used when the strip buffers fit in shared memory
  bool streaming = false;
  int sw = 128;           // strip width of the streaming form
  StreamArgs sa{};
  size_t ssmem = 0;
  dim3 sgrid;
  float* d_ws = nullptr;  // [wx' (WU + 8) | wy' (KW + 8)] x 2 orientations
  int swstride = 0;
  int map_rows() const { return streaming ? kCR : a.box_rows; }
  int map_cols() const { return streaming ? sa.bc : a.box_cols; }
};

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

// taps for conv #1 (u (*) psf) and conv #2 (r (*) mirror) as correlation taps
static int make_conv_plan(thz_ctx* c, cudaStream_t s, int Hp, int Wp, int pitch, const float* psf_x, int kx,
                          const float* psf_y, int ky, const float* dense, int direct, ConvPlan& cp) {
  if (kx < 1 || ky < 1 || (kx & 1) == 0 || (ky & 1) == 0) return set_err(c, THZ_EINVAL, "PSF extents must be odd");
  if (kx > THZ_MAX_PSF || ky > THZ_MAX_PSF) return set_err(c, THZ_EINVAL, "PSF larger than THZ_MAX_PSF");
  ConvArgs& a = cp.a;
  a.Hp = Hp; a.Wp = Wp; a.pitch = pitch; a.kx = kx; a.ky = ky;
  a.kxp = round_up(kx, 8);
  a.kyp = round_up(ky, 8);
  a.box_rows = kTH + kx - 1;
  int bc = kTW + a.kyp + 8;             // pass 1 reads up to column cg*8 + kyp + 15
  bc = round_up(bc, 4);
  if (((bc / 4) & 1) == 0) bc += 4;     // 4 * odd: conflict-free 128-bit row accesses
  a.box_cols = bc;
  a.eps = 1e-12f;
  a.col_shift = ((ky / 2) % 4 == 0) ? 0 : (ky / 2) % 4 - 4;
  cp.dense = dense != nullptr;
  if (a.box_rows > 256 || a.box_cols > 256) return set_err(c, THZ_EINVAL, "PSF too large for one TMA box");
  const int tile_floats = (a.box_rows * a.box_cols + 31) & ~31;
  std::vector<float> w;
  if (!cp.dense) {
    cp.smem = (size_t)(tile_floats + a.box_rows * kMidStride + a.kxp + a.kyp) * sizeof(float) + 256;
    cp.wstride = a.kxp + a.kyp;
    w.assign(2 * cp.wstride, 0.f);
    // orientation 0 = first conv of the iteration (u with psf), 1 = second (r with the mirrored psf).
    // direct branch: correlation with the given kernel; FFT branch: convolution = correlation with the flip.
    for (int o = 0; o < 2; ++o) {
      const bool flip = (o == 0) ? (direct == 0) : (direct != 0);
      for (int i = 0; i < kx; ++i) w[o * cp.wstride + i] = psf_x[flip ? kx - 1 - i : i];
      for (int j = 0; j < ky; ++j) w[o * cp.wstride + a.kxp + j] = psf_y[flip ? ky - 1 - j : j];
    }
  } else {
    cp.smem = (size_t)(tile_floats + kx * a.kyp) * sizeof(float) + 256;
    cp.wstride = kx * a.kyp;
    w.assign(2 * cp.wstride, 0.f);
    for (int o = 0; o < 2; ++o) {
      const bool flip = (o == 0) ? (direct == 0) : (direct != 0);
      for (int i = 0; i < kx; ++i)
        for (int j = 0; j < ky; ++j)
          w[o * cp.wstride + i * a.kyp + j] = dense[(flip ? kx - 1 - i : i) * ky + (flip ? ky - 1 - j : j)];
    }
  }
  if (cp.smem > 227 * 1024) return set_err(c, THZ_EINVAL, "PSF too large for the shared-memory tile");
  const size_t w_tile = w.size();
  if (!cp.dense) {
    StreamArgs& sa = cp.sa;
    sa.Hp = Hp; sa.Wp = Wp; sa.pitch = pitch;
    sa.WU = round_up(kx - 1, 8);
    sa.KW = round_up(ky - 1, 8);
    const int padx = sa.WU - (kx - 1), pady = sa.KW - (ky - 1);
    sa.gy_off = kx / 2 + padx;
    sa.gx_off = ky / 2 + pady;
    sa.Rg = 2 * kCR + sa.WU;
    sa.RS = sa.Rg + 4;
    if (((sa.RS / 4) & 1) == 0) sa.RS += 4;
    sa.col_shift = (sa.gx_off % 4 == 0) ? 0 : (sa.gx_off % 4) - 4;
    sa.eps = a.eps;
    auto box_cols = [&](int sw) {
      int v = round_up(sw + sa.KW, 4);
      if (((v / 4) & 1) == 0) v += 4;
      return v;
    };
    auto smem_of = [&](int sw) {
      return (size_t)(2 * kCR * box_cols(sw) + sw * sa.RS + sa.WU + 8 + sa.KW + 8) * sizeof(float) + 256;
    };
    // two 64-column CTAs per SM when both fit (each CTA also pays 1 KB of driver-reserved shared memory)
    cp.sw = (2 * (smem_of(64) + 1024) <= 228 * 1024) ? 64 : 128;
    sa.bc = box_cols(cp.sw);
    cp.ssmem = smem_of(cp.sw);
    cp.streaming = cp.ssmem <= 227 * 1024 && sa.bc <= 256;
    if (cp.streaming) {
      // segments: as many per strip as fill the resident CTA slots once, each a whole number of chunks
      const int strips = (Wp - sa.col_shift + cp.sw - 1) / cp.sw;
      const int slots = c->sm_count * (cp.sw == 64 ? 2 : 1);
      int segs = std::max(1, slots / strips);
      segs = std::min(segs, (Hp + kCR - 1) / kCR);
      const int per_seg = (Hp + segs - 1) / segs;
      const int chunks = (per_seg + sa.WU + kCR - 1) / kCR;
      sa.seg_rows = chunks * kCR - sa.WU;
      segs = (Hp + sa.seg_rows - 1) / sa.seg_rows;
      cp.sgrid = dim3(strips, segs);
      cp.swstride = sa.WU + 8 + sa.KW + 8;
      w.resize(w_tile + 2 * cp.swstride, 0.f);
      for (int o = 0; o < 2; ++o) {
        const bool flip = (o == 0) ? (direct == 0) : (direct != 0);
        float* wx = w.data() + w_tile + o * cp.swstride;
        float* wy = wx + sa.WU + 8;
        for (int i = 0; i < kx; ++i) wx[padx + i] = psf_x[flip ? kx - 1 - i : i];
        for (int j = 0; j < ky; ++j) wy[pady + j] = psf_y[flip ? ky - 1 - j : j];
      }
    }
  }
  void* dp = nullptr;
  int rc = ws_get(c, WS_RL_TAPS, w.size() * sizeof(float), &dp);
  if (rc != THZ_OK) return rc;
  cp.d_w = (float*)dp;
  cp.d_ws = cp.d_w + w_tile;
  THZ_CUDA(c, cudaStreamSynchronize(s));   // the previous band's kernels are done with the taps
  THZ_CUDA(c, cudaMemcpyAsync(cp.d_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice, s));
  THZ_CUDA(c, cudaStreamSynchronize(s));
  return THZ_OK;
}

template <int MODE>
static int launch_conv(thz_ctx* c, cudaStream_t s, const ConvPlan& cp, const CUtensorMap& map, int orient,
                       const float* d, float* out) {
  ConvArgs a = cp.a;
  if (cp.dense) {
    a.wdense = cp.d_w + (size_t)orient * cp.wstride;
  } else {
    a.wx = cp.d_w + (size_t)orient * cp.wstride;
    a.wy = a.wx + a.kxp;
  }
  a.d = d;
  a.out = out;
  dim3 grid((a.Wp - a.col_shift + kTW - 1) / kTW, (a.Hp + kTH - 1) / kTH);
  cudaError_t e;
  if (cp.streaming) {
    StreamArgs sa = cp.sa;
    sa.wx = cp.d_ws + (size_t)orient * cp.swstride;
    sa.wy = sa.wx + sa.WU + 8;
    sa.d = d;
    sa.out = out;
    auto launch = [&](auto kernel, int threads) -> cudaError_t {
      const void* skey = (const void*)kernel;
      size_t& shave = c->smem_set[skey];
      if (shave < cp.ssmem) {
        cudaError_t e2 = cudaFuncSetAttribute(skey, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cp.ssmem);
        if (e2 != cudaSuccess) return e2;
        shave = cp.ssmem;
      }
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = cp.sgrid;
      cfg.blockDim = dim3(threads);
      cfg.dynamicSmemBytes = cp.ssmem;
      cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      return cudaLaunchKernelEx(&cfg, kernel, map, sa);
    };
    e = (cp.sw == 64) ? launch(k_rl_stream<MODE, 64>, 256) : launch(k_rl_stream<MODE, 128>, 512);
    c->launches++;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(c, e, "k_rl_stream launch");
    return THZ_OK;
  }
  if (!cp.dense) {
    // persistent, double-buffered form: two haloed tiles + the intermediate tile
    const int tile_floats = (a.box_rows * a.box_cols + 31) & ~31;
    const size_t smem = (size_t)(2 * tile_floats + a.box_rows * kMidStride + a.kxp + a.kyp) * sizeof(float) + 256;
    if (smem <= 227 * 1024) {
      const void* pkey = (const void*)k_rl_conv_persistent<MODE>;
      size_t& phave = c->smem_set[pkey];
      if (phave < smem) {
        e = cudaFuncSetAttribute(pkey, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(c, e, "cudaFuncSetAttribute(rl persistent)");
        phave = smem;
      }
      const int ntiles = (int)(grid.x * grid.y);
      const int nb = std::min(ntiles, c->sm_count);
      k_rl_conv_persistent<MODE><<<nb, kRlThreads, smem, s>>>(map, a, (int)grid.x, (int)grid.y);
      c->launches++;
      e = cudaGetLastError();
      if (e != cudaSuccess) return cuda_fail(c, e, "k_rl_conv_persistent launch");
      return THZ_OK;
    }
  }
  const void* key = cp.dense ? (const void*)k_rl_conv<MODE, true> : (const void*)k_rl_conv<MODE, false>;
  size_t& have = c->smem_set[key];
  if (have < cp.smem) {
    e = cudaFuncSetAttribute(key, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cp.smem);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaFuncSetAttribute(rl)");
    have = cp.smem;
  }
  if (cp.dense) k_rl_conv<MODE, true><<<grid, 256, cp.smem, s>>>(map, a);
  else k_rl_conv<MODE, false><<<grid, 256, cp.smem, s>>>(map, a);
  c->launches++;
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "k_rl_conv launch");
  return THZ_OK;
}

static int grid_for(thz_ctx* c, int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)c->sm_count * 8;
  return (int)std::max<int64_t>(1, std::min(b, cap));
}

int conv2d_once(thz_ctx* c, cudaStream_t s, const float* d_in, int rows, int cols, const float* psf_x, int kx,
                const float* psf_y, int ky, const float* dense, int direct, float* d_out) {
  if (!d_in || !d_out || rows < 1 || cols < 1) return set_err(c, THZ_EINVAL, "bad image");
  const int pitch = round_up(cols, 4);
  void *pa = nullptr, *pb = nullptr;
  int rcw = ws_get(c, WS_CONV_A, (size_t)rows * pitch * sizeof(float), &pa);
  if (rcw == THZ_OK) rcw = ws_get(c, WS_CONV_B, (size_t)rows * pitch * sizeof(float), &pb);
  if (rcw != THZ_OK) return rcw;
  float *d_a = (float*)pa, *d_b = (float*)pb;
  k_copy2d<<<grid_for(c, (int64_t)rows * cols), 256, 0, s>>>(d_in, rows, cols, cols, d_a, pitch);
  c->launches++;
  ConvPlan cp;
  int rc = make_conv_plan(c, s, rows, cols, pitch, psf_x, kx, psf_y, ky, dense, direct ? 1 : 0, cp);
  CUtensorMap map;
  if (rc == THZ_OK) rc = make_tmap(c, &map, d_a, rows, cols, pitch, cp.map_rows(), cp.map_cols());
  // orientation 0 is "the kernel as given": correlation when direct, convolution otherwise
  if (rc == THZ_OK) rc = launch_conv<0>(c, s, cp, map, 0, nullptr, d_b);
  if (rc == THZ_OK) {
    k_copy2d<<<grid_for(c, (int64_t)rows * cols), 256, 0, s>>>(d_b, rows, cols, pitch, d_out, cols);
    c->launches++;
  }
  cudaStreamSynchronize(s);
  cudaError_t e = cudaGetLastError();
  if (rc == THZ_OK && e != cudaSuccess) return cuda_fail(c, e, "conv2d");
  return rc;
}

int richardson_lucy(thz_ctx* c, cudaStream_t s, const float* d_image, int rows, int cols, const float* psf_x, int kx,
                    const float* psf_y, int ky, const float* dense, int direct, int n_iter, float* d_deconv,
                    float* d_gain, const volatile uint8_t* abort_flag, thz_progress_fn progress, void* puser,
                    float pbase, float pspan) {
  if (!d_image || rows < 2 || cols < 2) return set_err(c, THZ_EINVAL, "bad image");
  const int pad_y = kx / 2, pad_x = ky / 2;   // psf.nrows()/2 pads axis 0 (deconvolution.rs:629-631)
  if (pad_y >= rows - 1 || pad_x >= cols - 1) return set_err(c, THZ_EINVAL, "PSF larger than the image");
  const int Hp = rows + 2 * pad_y, Wp = cols + 2 * pad_x, pitch = round_up(Wp, 4);
  const size_t img_bytes = (size_t)Hp * pitch * sizeof(float);
  void *pd = nullptr, *pu = nullptr, *pr = nullptr;
  int rcw = ws_get(c, WS_RL_D, img_bytes, &pd);
  if (rcw == THZ_OK) rcw = ws_get(c, WS_RL_U, img_bytes, &pu);
  if (rcw == THZ_OK) rcw = ws_get(c, WS_RL_R, img_bytes, &pr);
  if (rcw != THZ_OK) return rcw;
  float *d_d = (float*)pd, *d_u = (float*)pu, *d_r = (float*)pr;
  THZ_CUDA(c, cudaMemsetAsync(d_d, 0, img_bytes, s));
  k_reflect_pad<<<grid_for(c, (int64_t)Hp * Wp), 256, 0, s>>>(d_image, rows, cols, pad_y, pad_x, d_d, pitch);
  c->launches++;
  THZ_CUDA(c, cudaMemcpyAsync(d_u, d_d, img_bytes, cudaMemcpyDeviceToDevice, s));
  THZ_CUDA(c, cudaMemsetAsync(d_r, 0, img_bytes, s));
  ConvPlan cp;
  int rc = make_conv_plan(c, s, Hp, Wp, pitch, psf_x, kx, psf_y, ky, dense, direct ? 1 : 0, cp);
  CUtensorMap map_u, map_r;
  if (rc == THZ_OK) rc = make_tmap(c, &map_u, d_u, Hp, Wp, pitch, cp.map_rows(), cp.map_cols());
  if (rc == THZ_OK) rc = make_tmap(c, &map_r, d_r, Hp, Wp, pitch, cp.map_rows(), cp.map_cols());
  bool aborted = false;
  for (int it = 0; rc == THZ_OK && it < n_iter; ++it) {
    rc = launch_conv<1>(c, s, cp, map_u, 0, d_d, d_r);          // r = d / (u (*) psf + eps)
    if (rc == THZ_OK) rc = launch_conv<2>(c, s, cp, map_r, 1, nullptr, d_u);   // u *= r (*) mirror
    if ((it & 15) == 15 || it == n_iter - 1) {
      if (abort_flag && *abort_flag) { aborted = true; break; }
      if (progress || abort_flag) {
        cudaError_t e = cudaStreamSynchronize(s);   // keep the queue short so that abort is responsive
        if (e != cudaSuccess) { rc = cuda_fail(c, e, "richardson_lucy"); break; }
        if (progress) progress(pbase + pspan * (float)(it + 1) / (float)n_iter, puser);
      }
    }
  }
  if (rc == THZ_OK && !aborted) {
    k_rl_finish<<<grid_for(c, (int64_t)rows * cols), 256, 0, s>>>(d_u, pitch, pad_y, pad_x, rows, cols, d_image,
                                                                  d_deconv, d_gain);
    c->launches++;
  }
  cudaError_t e = cudaStreamSynchronize(s);
  if (rc == THZ_OK && e != cudaSuccess) return cuda_fail(c, e, "richardson_lucy");
  if (rc == THZ_OK && aborted) return THZ_ABORTED;
  return rc;
}

}  // namespace thz

using namespace thz;

#define CHECK_CTX(c)                                                   \
  do {                                                                 \
    if (!(c)) return THZ_EINVAL;                                       \
    cudaError_t e_ = cudaSetDevice((c)->device);                       \
    if (e_ != cudaSuccess) return cuda_fail((c), e_, "cudaSetDevice"); \
  } while (0)

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

int thz_rl_separable_dev(thz_ctx* c, const float* d_image, int rows, int cols, const float* psf_x, int kx,
                         const float* psf_y, int ky, int direct, int n_iter, float* d_deconvolved, float* d_gain,
                         const volatile uint8_t* abort_flag, thz_progress_fn progress, void* progress_user,
                         float progress_base, float progress_span) {
  CHECK_CTX(c);
  if (!psf_x || !psf_y) return set_err(c, THZ_EINVAL, "null PSF");
  return richardson_lucy(c, c->stream, d_image, rows, cols, psf_x, kx, psf_y, ky, nullptr, direct, n_iter,
                         d_deconvolved, d_gain, abort_flag, progress, progress_user, progress_base, progress_span);
}

int thz_rl_dense_dev(thz_ctx* c, const float* d_image, int rows, int cols, const float* psf, int kx, int ky, int direct,
                     int n_iter, float* d_deconvolved, float* d_gain, const volatile uint8_t* abort_flag) {
  CHECK_CTX(c);
  if (!psf) return set_err(c, THZ_EINVAL, "null PSF");
  return richardson_lucy(c, c->stream, d_image, rows, cols, nullptr, kx, nullptr, ky, psf, direct, n_iter,
                         d_deconvolved, d_gain, abort_flag, nullptr, nullptr, 0.f, 0.f);
}

int thz_conv2d_separable_dev(thz_ctx* c, const float* d_in, int rows, int cols, const float* psf_x, int kx,
                             const float* psf_y, int ky, int direct, float* d_out) {
  CHECK_CTX(c);
  if (!psf_x || !psf_y) return set_err(c, THZ_EINVAL, "null PSF");
  return conv2d_once(c, c->stream, d_in, rows, cols, psf_x, kx, psf_y, ky, nullptr, direct, d_out);
}

int thz_conv2d_dense_dev(thz_ctx* c, const float* d_in, int rows, int cols, const float* psf, int kx, int ky,
                         int direct, float* d_out) {
  CHECK_CTX(c);
  if (!psf) return set_err(c, THZ_EINVAL, "null PSF");
  return conv2d_once(c, c->stream, d_in, rows, cols, nullptr, kx, nullptr, ky, psf, direct, d_out);
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
  for (int b = 0; rc == THZ_OK && b < n_bands; ++b) {
    if (abort_flag && *abort_flag) { rc = THZ_ABORTED; break; }
    const float base = 0.1f + 0.8f * (float)done_iter / (float)total_iter;
    const float span = 0.8f * (float)std::max(bands[b].n_iter, 1) / (float)total_iter;
    rc = richardson_lucy(c, c->stream, d_energy + (size_t)b * P, rows, cols, bands[b].psf_x, bands[b].kx,
                         bands[b].psf_y, bands[b].ky, nullptr, bands[b].direct, bands[b].n_iter, nullptr,
                         d_gain + (size_t)b * P, abort_flag, progress, progress_user, base, span);
    done_iter += std::max(bands[b].n_iter, 1);
  }
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
  if (c->plan.n != n) return set_err(c, THZ_ESTATE, "thz_plan_trace(n, ...) must be called first");
  const int64_t P = (int64_t)rows * cols;
  if (P == 0) return THZ_OK;
  if (!cube || !out) return set_err(c, THZ_EINVAL, "null pointer");
  if (n_bands < 0 || n_bands > THZ_MAX_BANDS || (n_bands > 0 && !bands)) return set_err(c, THZ_EINVAL, "bad bands");
  void *pc = nullptr, *pi = nullptr, *pe = nullptr, *pg = nullptr;
  int rc = ws_get(c, WS_HOST_CUBE, (size_t)P * n * sizeof(float), &pc);
  if (rc == THZ_OK) rc = ws_get(c, WS_HOST_IMG, (size_t)P * sizeof(float), &pi);
  if (rc == THZ_OK && n_bands) rc = ws_get(c, WS_ENERGY, (size_t)n_bands * P * sizeof(float), &pe);
  if (rc == THZ_OK && n_bands) rc = ws_get(c, WS_GAIN, (size_t)n_bands * P * sizeof(float), &pg);
  if (rc != THZ_OK) return rc;
  float *d_cube = (float*)pc, *d_img = (float*)pi, *d_energy = (float*)pe, *d_gain = (float*)pg;
  if (progress) progress(0.0f, progress_user);
  // FIR tables are built (and cached) before the pipelined loop so that no chunk waits on the host
  if (n_bands) {
    FirTables ft;
    rc = upload_fir_tables(c, c->stream, n, bands, n_bands, ft);
    if (rc != THZ_OK) return rc;
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
    rc = launch_trace_fused(c, s, d, d, d_img + p, np);
    if (rc == THZ_OK && n_bands) rc = deconv_energies(c, s, d, np, n, bands, n_bands, d_energy + p, P);
    if (rc == THZ_OK && !n_bands) {
      THZ_CUDA(c, cudaMemcpyAsync(out + p * n, d, (size_t)np * n * sizeof(float), cudaMemcpyDeviceToHost, s));
    }
  }
  for (int k = 0; k < kHostStreams; ++k) THZ_CUDA(c, cudaStreamSynchronize(c->hstream[k]));
  if (rc != THZ_OK) return rc;
  if (n_bands) {
    long total_iter = 0, done_iter = 0;
    for (int b = 0; b < n_bands; ++b) total_iter += std::max(bands[b].n_iter, 1);
    for (int b = 0; rc == THZ_OK && b < n_bands; ++b) {
      if (abort_flag && *abort_flag) return THZ_ABORTED;
      const float base = 0.1f + 0.8f * (float)done_iter / (float)total_iter;
      const float span = 0.8f * (float)std::max(bands[b].n_iter, 1) / (float)total_iter;
      rc = richardson_lucy(c, c->stream, d_energy + (size_t)b * P, rows, cols, bands[b].psf_x, bands[b].kx,
                           bands[b].psf_y, bands[b].ky, nullptr, bands[b].direct, bands[b].n_iter, nullptr,
                           d_gain + (size_t)b * P, abort_flag, progress, progress_user, base, span);
      done_iter += std::max(bands[b].n_iter, 1);
    }
    if (rc != THZ_OK) return rc;
    i = 0;
    for (int64_t p = 0; p < P && rc == THZ_OK; p += ct, ++i) {
      const int64_t np = std::min(ct, P - p);
      cudaStream_t s = c->hstream[i % kHostStreams];
      float* d = d_cube + p * n;
      rc = deconv_apply(c, s, d, d_gain + p, np, n, bands, n_bands, d, d_img + p, P, 1 + i % kHostStreams);
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
  if (progress) progress(1.0f, progress_user);
  return THZ_OK;
}

}  // extern "C"
