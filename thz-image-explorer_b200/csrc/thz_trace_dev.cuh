// thz_trace_dev.cuh -- device-side pieces of the trace pass shared by thz_trace.cu and the fused
// trace + band-energy kernel in thz_deconv.cu: CTA geometry, kernel arguments, the pair load with the
// time-domain multiplier and the pair store with the gate, the zero-trace rule and the intensity sum.
#pragma once
#include "thz_fft.cuh"
#include "thz_internal.h"

namespace thz {

template <int N> struct Geo {
  static constexpr int T = N / kE;                       // threads per trace pair
  static constexpr int NT = (T >= 256) ? T : 256;        // threads per CTA
  static constexpr int G = NT / T;                       // trace pairs per CTA pass
  static constexpr int kScr = (32 + kNzWords) * G;       // per-CTA scratch words: reductions + zero-trace flags
  static constexpr int kMinBlocks = (NT == 256) ? 2 : 1; // register cap: 128 per thread
  static constexpr size_t smem_bytes = (size_t)G * padded_len(N) * sizeof(float2) + kScr * sizeof(float);
  // kernels that stage their input with bulk copies: two slabs of 2*G traces + two mbarriers
  static constexpr int kSlabFloats = 2 * G * N;
  static constexpr size_t stage_off = (smem_bytes + 127) & ~(size_t)127;
  static constexpr size_t smem_bytes_staged = stage_off + 2 * (size_t)kSlabFloats * sizeof(float) + 16;
};

struct TraceArgs {
  const float* in;        // [P][N]
  float* out;             // [P][N]
  float* img;             // [P] or null
  const float* m_pre;     // [N] or null
  const float* m_post;    // [N] or null
  int pre_ends, post_ends; // multiplier is exactly 1 away from the first / last N/16 samples (windows and gates
                           // usually are): multiplying by 1.0f is the identity, so only registers 0 and kE-1 load it
  const float* hq;        // [N] (band / N) in last-stage register order
  const float* band;      // [F] or null
  const float2* tw;
  float2* fft;            // [P][F]
  const float2* fft_in;   // [P][F]
  float* amp;
  float* phase;
  float* win;             // windowed trace out (forward) or null
  // reference-pulse normalisation of the forward outputs (config 2): amp -> A_s / max(A_r, 1e-12),
  // phase -> phi_s - phi_r (the operands of calculate_optical_properties, src/math_tools.rs:665-701); null = off
  const float* ref_amp;   // [F]
  const float* ref_phase; // [F]
  int64_t P;
};

__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }

// sum `a` and `b` over the T threads of a group; result valid in thread t == 0 of the group
template <int N>
__device__ __forceinline__ void group_reduce2(float& a, float& b, int t, int g, float* scr) {
  constexpr int T = Geo<N>::T;
  constexpr int W = (T < 32) ? T : 32;
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if constexpr (T > 32) {
    constexpr int NW = T / 32;
    float* s = scr + g * 32;
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

// The two halves of load_pair, for kernels that fetch the NEXT pair while they finish the current one: the loads
// alone (nothing depends on them, their L2 / DRAM latency hides behind the rest of the item), then the zero test
// and the time-domain multiplier.
template <int N>
__device__ __forceinline__ void load_pair_raw(float2 (&v)[kE], const float* __restrict__ in, int t, bool act0, bool act1,
                                              int64_t p0) {
  constexpr int T = Geo<N>::T;
  const float* r0 = in + p0 * N + t;
  const float* r1 = r0 + N;
#pragma unroll
  for (int i = 0; i < kE; ++i) {
    v[i].x = act0 ? ld_stream(r0 + i * T) : 0.f;
    v[i].y = act1 ? ld_stream(r1 + i * T) : 0.f;
  }
}

template <int N>
__device__ __forceinline__ void finish_pair(float2 (&v)[kE], const TraceArgs& a, int t, bool& nz0, bool& nz1) {
  constexpr int T = Geo<N>::T;
  nz0 = nz1 = false;
#pragma unroll
  for (int i = 0; i < kE; ++i) {
    nz0 |= (v[i].x != 0.f);
    nz1 |= (v[i].y != 0.f);
  }
  if (a.m_pre != nullptr) {
    if (a.pre_ends) {
#pragma unroll
      for (int i = 0; i < kE; i += kE - 1) {
        const float m = __ldg(a.m_pre + t + i * T);
        v[i].x *= m;
        v[i].y *= m;
      }
    } else {
#pragma unroll
      for (int i = 0; i < kE; ++i) {
        const float m = __ldg(a.m_pre + t + i * T);
        v[i].x *= m;
        v[i].y *= m;
      }
    }
  }
}

// Load the pair (p0, p1) into stage-0 register layout, multiplied by m_pre.
template <int N>
__device__ __forceinline__ void load_pair(float2 (&v)[kE], const TraceArgs& a, int t, bool act0, bool act1,
                                          int64_t p0, bool& nz0, bool& nz1) {
  load_pair_raw<N>(v, a.in, t, act0, act1, p0);
  finish_pair<N>(v, a, t, nz0, nz1);
}

// multiply by m_post, store both traces, intensity = sum of squares of the stored values
// (CUBE = false: everything but the stores of the traces -- the caller keeps them on chip)
template <int N, bool CUBE = true>
__device__ __forceinline__ void store_pair(float2 (&v)[kE], const TraceArgs& a, int t, int g, bool act0,
                                           bool act1, int64_t p0, bool use_post, float* scr, bool z0, bool z1) {
  constexpr int T = Geo<N>::T;
  if (z0 || z1) {   // all-zero input trace: exact zeros out, as when transformed on its own
#pragma unroll
    for (int i = 0; i < kE; ++i) {
      if (z0) v[i].x = 0.f;
      if (z1) v[i].y = 0.f;
    }
  }
  if (use_post) {
    if (a.post_ends) {
#pragma unroll
      for (int i = 0; i < kE; i += kE - 1) {
        const float m = __ldg(a.m_post + t + i * T);
        v[i].x *= m;
        v[i].y *= m;
      }
    } else {
#pragma unroll
      for (int i = 0; i < kE; ++i) {
        const float m = __ldg(a.m_post + t + i * T);
        v[i].x *= m;
        v[i].y *= m;
      }
    }
  }
  float* r0 = a.out + p0 * N + t;
  float* r1 = r0 + N;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < kE; ++i) {
    if constexpr (CUBE) {
      if (act0) st_stream(r0 + i * T, v[i].x);
      if (act1) st_stream(r1 + i * T, v[i].y);
    }
    s0 = fmaf(v[i].x, v[i].x, s0);
    s1 = fmaf(v[i].y, v[i].y, s1);
  }
  if (a.img != nullptr) {       // uniform over the CTA
    group_reduce2<N>(s0, s1, t, g, scr);
    if (t == 0) {
      if (act0) a.img[p0] = s0;
      if (act1) a.img[p0 + 1] = s1;
    }
  }
}

// thz_trace.cu: arguments of the fused trace pass from the context's plan (tables, multipliers, band)
int trace_fused_args(thz_ctx* c, TraceArgs& a, const float* d_in, float* d_out, float* d_img, int64_t P);

}  // namespace thz
