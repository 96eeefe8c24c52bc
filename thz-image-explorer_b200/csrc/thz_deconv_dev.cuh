// thz_deconv_dev.cuh -- device code shared by the two translation units of the FIR passes: thz_deconv.cu (passes A
// and C, edge kernels, host side) and thz_chain_fused.cu (the fused trace + band-energy kernel, compiled on its own
// so that the two large sets of template instantiations build in parallel).
#pragma once
#include "thz_trace_dev.cuh"

#include <math.h>

namespace thz {

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
  // spectral hand-off between the fused trace + energy kernel and pass C (see k_chain_energy_fused)
  float* edges;          // [P][512]: samples [0, 256) and [N - 256, N) of every filtered trace
  const float2* xspec;   // [P / 2][N]: FFT_N of the filtered pair in Plan<N> register order (pass C input)
};

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

// Band energies of one pair from the Parseval terms of the two sub-spectra: e[band][trace] = sum over the
// lower-half registers of we q_even + wo q_odd (+ the Nyquist bin), reduced over the group and written to
// energy[band][p0 .. p0+1].  `red` is group-private shared memory nobody else reads any more.
template <int N>
__device__ __forceinline__ void band_energy_reduce(const FirArgs& a, const float (&q1e)[kE / 2], const float (&q2e)[kE / 2],
                                                   const float (&q1o)[kE / 2], const float (&q2o)[kE / 2], float ny1,
                                                   float ny2, int t, float* red, int64_t p0, bool act0, bool act1,
                                                   bool z0, bool z1) {
  constexpr int T = SGeo<N>::T;
  constexpr int LAST = Plan<N>::ns - 1;
  constexpr int RL = Plan<N>::r[LAST];
  constexpr int UL = kE / RL;
  constexpr int NLOW = kE / 2;
  constexpr int W = (T < 32) ? T : 32;
  constexpr int NW = (T + 31) / 32;
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

// ------------------------------------------------------------------------------------
// Trace pass + pass A in ONE kernel: window -> FFT -> band-pass -> inverse FFT -> gate -> store +
// intensity, and, while the filtered pair is still on chip, the Parseval band energies of the full linear FIR
// convolution (the edge kernel subtracts the cut-off samples afterwards, as after k_fir_energy_split).
//
// Odd bins of the M = 2N spectrum: the gated pair y is in registers after the store -> modulate -> one forward
// transform.  Even bins are FFT_N(y).  POST selects how they are obtained:
//   0  no gate after the inverse: FFT_N(y) = N Z' with Z' = (band / N) X, the filtered spectrum this kernel held
//      before the inverse transform (kept in a thread-private shared-memory stash) -- no transform at all;
//   1  the gate differs from 1 only in the first / last four samples (the default gate: a 0.1 ps edge at 0.05 ps
//      steps): y = y0 - c with c supported on those eight samples, so FFT_N(y)[k] = N Z'[k] - sum_n c_n w_N^(k n).
//      A thread's 16 bins are k0(u) + m N/RL (m = digit of the last radix-RL stage), hence
//      w_N^(k n) = w_N^(k0 n) exp(-2 pi i m n / RL): the correction of its RL bins is ONE RL-point DFT of the eight
//      twiddled samples -- a few hundred flops instead of a 4096-point transform;
//   2  general gate: the stash holds y instead and the even bins cost a second forward transform.
// Modes 0 / 1 differ from transforming the stored trace by f32 rounding only (tests/test_chain_fused_gpu.py).
//
// SPEC (spectral hand-off, whole-chain calls): the kernel writes FFT_N of the gated pair -- which it holds anyway
// -- INSTEAD of the filtered traces, same bytes, same place; pass C (k_fir_apply_circ<.., SPEC>) then starts from
// the spectrum and saves its forward transform.  The 2 x 256 edge samples per trace that k_fir_edges(_mma) and
// k_fir_edge_corr read go to a side buffer of 512-float rows (those kernels see it as a cube of 512-sample traces,
// of which they only touch the first and last 249 / 256 samples).  Trace counts must be even (whole pairs).
// ------------------------------------------------------------------------------------
template <int N> struct FGeo {
  static constexpr size_t stash_off = (Geo<N>::smem_bytes + 15) & ~(size_t)15;
  static constexpr size_t smem_bytes = stash_off + (size_t)Geo<N>::G * (N + 8) * sizeof(float2);
};

template <int N, int POST, bool SPEC>
__global__ void __launch_bounds__(Geo<N>::NT, Geo<N>::kMinBlocks) k_chain_energy_fused(const TraceArgs a, const FirArgs f) {
  using GEO = Geo<N>;
  constexpr int T = GEO::T, G = GEO::G;
  static_assert(T >= 8, "head and tail gate samples must belong to different threads");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  float* scr = reinterpret_cast<float*>(smem + (size_t)G * padded_len(N));
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  float2* sm = smem + (size_t)g * padded_len(N);
  float2* stash = reinterpret_cast<float2*>(smem_raw + FGeo<N>::stash_off) + (size_t)g * N + t;   // [i * T]: thread-private
  float2* cs = reinterpret_cast<float2*>(smem_raw + FGeo<N>::stash_off) + (size_t)G * N + g * 8;  // gate corrections of the group
  const int64_t npairs = (a.P + 1) >> 1;
  const int64_t nitems = (npairs + G - 1) / G;
  constexpr int LAST = Plan<N>::ns - 1;
  constexpr int RL = Plan<N>::r[LAST];
  constexpr int UL = kE / RL;
  constexpr int NLOW = kE / 2;
  unsigned* nzbuf = reinterpret_cast<unsigned*>(scr + 32 * G);
  float* red = reinterpret_cast<float*>(sm);
  int parity = 0;

  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x, parity ^= 1) {
    const int64_t p0 = (item * G + g) * 2;
    const bool act0 = p0 < a.P, act1 = p0 + 1 < a.P;
    const int64_t next = item + gridDim.x;
    if (next < nitems) {
      int64_t cnt = a.P - next * G * 2;
      if (cnt > 2 * G) cnt = 2 * G;
      prefetch_l2_slab(a.in + next * G * 2 * N, cnt * N);
    }
    // (fetching the next pair into registers across the band sums, as k_trace_fused does across its stores, costs
    // 1.5 % here: measured, gpurun call R of round 2)
    float2 v[kE];
    bool nz0, nz1, z0, z1;
    load_pair<N>(v, a, t, act0, act1, p0, nz0, nz1);
    nz_publish<T>(nz0, nz1, t, g, parity, nzbuf, z0, z1);
    {
      float hq[kE];
      auto fetch_hq = [&]() {
#pragma unroll
        for (int i = 0; i < kE; ++i) {
          const int u = i % UL, m = i / UL;
          hq[i] = __ldg(a.hq + m * (N / RL) + t + u * T);
        }
      };
      fft_forward_hook<N>(v, t, sm, a.tw, fetch_hq);
#pragma unroll
      for (int i = 0; i < kE; ++i) {
        v[i].x *= hq[i];
        v[i].y *= hq[i];
      }
    }
    if constexpr (POST != 2) {
#pragma unroll
      for (int i = 0; i < kE; ++i) stash[i * T] = v[i];
    }
    fft_inverse<N>(v, t, sm, a.tw);
    nz_resolve<T>(g, parity, nzbuf, z0, z1);
    float2 y0h = v[0], y0t = v[kE - 1];   // samples t and N - T + t before the gate
    store_pair<N, !SPEC>(v, a, t, g, act0, act1, p0, a.m_post != nullptr, scr, z0, z1);
    if constexpr (SPEC) {
      // the filtered traces themselves are not written: pass C starts from their spectrum (stored below) and the
      // edge kernels from the first / last 256 samples, kept as rows of 512 floats
      float* e0 = f.edges + p0 * 512;
      float* e1 = e0 + 512;
      constexpr int NE = (T >= 256) ? 1 : 256 / T;   // registers that hold head (tail) samples
#pragma unroll
      for (int i = 0; i < NE; ++i) {
        const int n = t + i * T;
        if (T <= 256 || n < 256) {
          if (act0) e0[n] = v[i].x;
          if (act1) e1[n] = v[i].y;
        }
      }
#pragma unroll
      for (int i = kE - NE; i < kE; ++i) {
        const int n = t + i * T - (N - 512);           // 256 + (sample - (N - 256))
        if (T <= 256 || n >= 256) {
          if (act0) e0[n] = v[i].x;
          if (act1) e1[n] = v[i].y;
        }
      }
    }
    if constexpr (POST == 1) {
      // c_n = y0[n] - y[n] for n = 0..3 (cs[n]) and n = N - j, j = 1..4 (cs[3 + j]); read after the barriers of the
      // transform below, overwritten only after the barriers of the next item
      if (t < 4) cs[t] = csub(y0h, v[0]);
      if (t >= T - 4) cs[3 + (T - t)] = csub(y0t, v[kE - 1]);
    } else {
      (void)y0h;
      (void)y0t;
    }
    if constexpr (POST == 2) {
#pragma unroll
      for (int i = 0; i < kE; ++i) stash[i * T] = v[i];
    }
    // ---- odd bins: FFT_N(y w_M^n) ----
    float q1o[NLOW], q2o[NLOW];
    modulate<N>(v, f.mod, t);
    fft_forward<N>(v, t, sm, a.tw);
    __syncthreads();
#pragma unroll
    for (int i = kE / 2; i < kE; ++i) sm[pad_idx(pos_to_bin<N>(stage_elem<N, LAST>(t, i)))] = v[i];   // mirrors of lower-half bins are upper-half registers
    // POST != 2: the even bins need no exchange, so their mirrors go to the other half of the buffer (the mirrors of
    // either sub-spectrum are bins >= N/2) and ONE barrier serves both sets of partner reads
    if constexpr (POST == 2) {
      __syncthreads();
      parseval_terms<N, true>(v, sm, t, q1o, q2o);
    }
    // ---- even bins: FFT_N(y) ----
    float2 z[kE];
#pragma unroll
    for (int i = 0; i < kE; ++i) z[i] = stash[i * T];
    if constexpr (POST == 2) {
      fft_forward<N>(z, t, sm, a.tw);   // its first barrier orders the partner reads above before the exchange
    } else {
#pragma unroll
      for (int i = 0; i < kE; ++i) {
        z[i].x *= (float)N;
        z[i].y *= (float)N;
      }
      if constexpr (POST == 1) {
        float2 c[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = cs[j];
#pragma unroll
        for (int u = 0; u < UL; ++u) {
          const int k0 = pos_to_bin<N>(stage_elem<N, LAST>(t, u));   // bin of digit m = 0; < N / RL
          const float2 w1 = __ldg(f.mod + 2 * k0);                   // exp(-2 pi i k0 / N)
          const float2 w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2);
          float2 d[RL];
#pragma unroll
          for (int m = 0; m < RL; ++m) d[m] = make_float2(0.f, 0.f);
          d[0] = c[0];
          d[1 % RL] = cadd(d[1 % RL], cmul(c[1], w1));
          d[2 % RL] = cadd(d[2 % RL], cmul(c[2], w2));
          d[3 % RL] = cadd(d[3 % RL], cmul(c[3], w3));
          d[(RL - 1) % RL] = cadd(d[(RL - 1) % RL], cmul_conj(c[4], w1));           // n = N - 1: w_N^(-k0)
          d[(2 * RL - 2) % RL] = cadd(d[(2 * RL - 2) % RL], cmul_conj(c[5], w2));
          d[(3 * RL - 3) % RL] = cadd(d[(3 * RL - 3) % RL], cmul_conj(c[6], w3));
          d[(4 * RL - 4) % RL] = cadd(d[(4 * RL - 4) % RL], cmul_conj(c[7], w4));
          dftR<RL, false>(d);
#pragma unroll
          for (int m = 0; m < RL; ++m) z[u + UL * m] = csub(z[u + UL * m], d[m]);
        }
      }
    }
    if constexpr (SPEC) {
      // FFT_N of the gated pair in register order: what pass C would otherwise recompute from the stored traces
      float2* so = reinterpret_cast<float2*>(a.out) + (p0 >> 1) * N + t;
      if (act0) {
#pragma unroll
        for (int i = 0; i < kE; ++i) __stcs(so + i * T, z[i]);
      }
    }
    float2* sme = (POST == 2) ? sm : sm - pad_idx(N / 2);   // pad_idx(i) - pad_idx(N/2) = pad_idx(i - N/2) for i >= N/2
    if constexpr (POST == 2) __syncthreads();   // the odd-bin partner reads and the last exchange of the transform are done
#pragma unroll
    for (int i = kE / 2; i < kE; ++i) sme[pad_idx(pos_to_bin<N>(stage_elem<N, LAST>(t, i)))] = z[i];
    __syncthreads();
    if constexpr (POST != 2) parseval_terms<N, true>(v, sm, t, q1o, q2o);
    float q1e[NLOW], q2e[NLOW];
    parseval_terms<N, false>(z, sme, t, q1e, q2e);
    float ny1 = 0.f, ny2 = 0.f;   // bin M/2 = even index N/2: register (u = 0, digit RL/2) of thread 0
    if (t == 0) {
      const float2 zz = z[UL * (RL / 2)];
      ny1 = zz.x * zz.x;
      ny2 = zz.y * zz.y;
    }
    __syncthreads();   // partner reads done: the buffer becomes reduction scratch
    band_energy_reduce<N>(f, q1e, q2e, q1o, q2o, ny1, ny2, t, red, p0, act0, act1, z0, z1);
  }
}

// thz_chain_fused.cu: launches k_chain_energy_fused<n, post, spectral hand-off when fa.edges != null>
int dispatch_chain_fused(thz_ctx* c, cudaStream_t s, int n, const TraceArgs& ta, const FirArgs& fa, int post);

}  // namespace thz
