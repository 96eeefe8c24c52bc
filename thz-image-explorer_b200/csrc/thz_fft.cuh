// thz_fft.cuh -- in-shared-memory mixed-radix complex FFT building blocks for sm_100a.
//
// One "group" of T = N/16 threads transforms one complex sequence of N points; every
// thread keeps E = 16 complex points in registers.  A forward transform is a chain of
// decimation-in-frequency (DIF) stages, each a radix-R (R in {2,4,8,16}) register
// butterfly followed by a twiddle multiply; stages are separated by one exchange through
// a padded float2 array in shared memory.  The result is left in *digit-reversed*
// position order.  The inverse transform is the exact conjugate transpose (DIT): it
// consumes digit-reversed order and produces natural order, so a forward -> pointwise
// multiply -> inverse pipeline (the filter chain) needs no permutation and no exchange
// between the last forward stage and the first inverse stage.
//
// This replaces the reference's per-trace calls into realfft/rustfft
// (src/math_tools.rs:374-375 forward, :559-566 inverse; src/filters/deconvolution.rs:266-317).
// Two real traces are packed as re/im of one complex sequence.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace thz {

constexpr int kE = 16;          // complex points per thread
constexpr int kPadShift = 4;    // one float2 of padding per 16 elements

__host__ __device__ constexpr int pad_idx(int i) { return i + (i >> kPadShift); }
__host__ __device__ constexpr int padded_len(int n) { return n + (n >> kPadShift); }

// Radix plans (stage 0 first).  Chosen with tools/bank_conflicts.py so that every exchange
// and the natural-order scatter are bank-conflict free (8192: at most 2-way on the scatter).
template <int N> struct Plan;
template <> struct Plan<64>   { static constexpr int ns = 2; static constexpr int r[4] = {4, 16, 1, 1}; };
template <> struct Plan<128>  { static constexpr int ns = 2; static constexpr int r[4] = {8, 16, 1, 1}; };
template <> struct Plan<256>  { static constexpr int ns = 2; static constexpr int r[4] = {16, 16, 1, 1}; };
template <> struct Plan<512>  { static constexpr int ns = 3; static constexpr int r[4] = {8, 16, 4, 1}; };
template <> struct Plan<1024> { static constexpr int ns = 3; static constexpr int r[4] = {16, 16, 4, 1}; };
template <> struct Plan<2048> { static constexpr int ns = 3; static constexpr int r[4] = {16, 16, 8, 1}; };
template <> struct Plan<4096> { static constexpr int ns = 3; static constexpr int r[4] = {16, 16, 16, 1}; };
template <> struct Plan<8192> { static constexpr int ns = 4; static constexpr int r[4] = {8, 4, 16, 16}; };

// sub-transform length at stage s: L_s = N / (r_0 ... r_{s-1})
template <int N, int S> struct StageLen {
  static constexpr int value = StageLen<N, S - 1>::value / Plan<N>::r[S - 1];
};
template <int N> struct StageLen<N, 0> { static constexpr int value = N; };

// Offset (in float2) of stage s's twiddle table inside the per-N table.  Stage s (all but
// the last) stores (R-1) * S entries laid out [q-1][j], j in [0, S), S = L/R:
//   tw[q-1][j] = exp(-2 pi i * j * q / L)
template <int N, int S> struct TwOffset {
  static constexpr int Lp = StageLen<N, S - 1>::value;
  static constexpr int Rp = Plan<N>::r[S - 1];
  static constexpr int value = TwOffset<N, S - 1>::value + (Rp - 1) * (Lp / Rp);
};
template <int N> struct TwOffset<N, 0> { static constexpr int value = 0; };
template <int N> struct TwTotal { static constexpr int value = TwOffset<N, Plan<N>::ns - 1>::value; };

// ------------------------------------------------------------------------------------
// complex helpers
// ------------------------------------------------------------------------------------
#ifndef THZ_SCALAR_ADD
// packed FP32 (add.rn.f32x2 / fma.rn.f32x2): one issue slot per complex add / subtract instead of two, same rounding
// per lane.  15 % fewer instructions in the transform kernels, 1-2 % less time (the FP32 pipe spends two cycles on a
// packed instruction; profiles/r02_ncu_full_chain_fused.txt).  -DTHZ_SCALAR_ADD builds the scalar form (A/B).
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 w) {  // a * conj(w)
  return make_float2(fmaf(a.x, w.x, a.y * w.y), fmaf(a.y, w.x, -a.x * w.y));
}
// multiply by -i (forward) or +i (inverse)
template <bool INV> __device__ __forceinline__ float2 rot90(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// multiply by exp(-+ i pi/4)
template <bool INV> __device__ __forceinline__ float2 rot45(float2 a) {
  constexpr float h = 0.70710678118654752440f;
  return INV ? make_float2((a.x - a.y) * h, (a.x + a.y) * h) : make_float2((a.x + a.y) * h, (a.y - a.x) * h);
}
// multiply by exp(-+ i 3pi/4)
template <bool INV> __device__ __forceinline__ float2 rot135(float2 a) {
  constexpr float h = 0.70710678118654752440f;
  return INV ? make_float2((-a.x - a.y) * h, (a.x - a.y) * h) : make_float2((a.y - a.x) * h, (-a.x - a.y) * h);
}
// multiply by exp(-+ 2 pi i k / 16) with compile-time (c, s) = (cos, sin)
template <bool INV> __device__ __forceinline__ float2 rotcs(float2 a, float c, float s) {
  // forward: (x + iy)(c - is) = (xc + ys) + i(yc - xs)
  return INV ? make_float2(fmaf(a.x, c, -a.y * s), fmaf(a.y, c, a.x * s))
             : make_float2(fmaf(a.x, c, a.y * s), fmaf(a.y, c, -a.x * s));
}

// ------------------------------------------------------------------------------------
// register butterflies: in-place DFT of R points, natural order in and out
// ------------------------------------------------------------------------------------
template <bool INV> __device__ __forceinline__ void dft2(float2& a0, float2& a1) {
  float2 t = a0;
  a0 = cadd(t, a1);
  a1 = csub(t, a1);
}

template <bool INV> __device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
  float2 t2 = cadd(a1, a3), t3 = rot90<INV>(csub(a1, a3));
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = cadd(t1, t3);
  a3 = csub(t1, t3);
}

template <bool INV> __device__ __forceinline__ void dft8(float2 (&a)[8]) {
  dft4<INV>(a[0], a[2], a[4], a[6]);   // even samples -> E[0..3] in a0,a2,a4,a6
  dft4<INV>(a[1], a[3], a[5], a[7]);   // odd samples  -> O[0..3] in a1,a3,a5,a7
  float2 o1 = rot45<INV>(a[3]), o2 = rot90<INV>(a[5]), o3 = rot135<INV>(a[7]);
  float2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], o0 = a[1];
  a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
  a[1] = cadd(e1, o1); a[5] = csub(e1, o1);
  a[2] = cadd(e2, o2); a[6] = csub(e2, o2);
  a[3] = cadd(e3, o3); a[7] = csub(e3, o3);
}

template <bool INV> __device__ __forceinline__ void dft16(float2 (&a)[16]) {
  constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
  // n = 4*a_ + b : inner DFT4 over a_ for each b -> Y_b[c] stored at a[4c + b]
#pragma unroll
  for (int b = 0; b < 4; ++b) dft4<INV>(a[b], a[4 + b], a[8 + b], a[12 + b]);
  // twiddle w16^(b*c)
  a[5] = rotcs<INV>(a[5], c1, s1);      // b=1,c=1 : k=1
  a[9] = rot45<INV>(a[9]);              // b=1,c=2 : k=2
  a[13] = rotcs<INV>(a[13], s1, c1);    // b=1,c=3 : k=3
  a[6] = rot45<INV>(a[6]);              // b=2,c=1 : k=2
  a[10] = rot90<INV>(a[10]);            // b=2,c=2 : k=4
  a[14] = rot135<INV>(a[14]);           // b=2,c=3 : k=6
  a[7] = rotcs<INV>(a[7], s1, c1);      // b=3,c=1 : k=3
  a[11] = rot135<INV>(a[11]);           // b=3,c=2 : k=6
  a[15] = rotcs<INV>(a[15], -c1, -s1);  // b=3,c=3 : k=9 -> cos = -c1, sin(2pi9/16) = -s1
  // outer DFT4 over b for each c -> X[c + 4d] lands in a[4c + d]
#pragma unroll
  for (int c = 0; c < 4; ++c) dft4<INV>(a[4 * c], a[4 * c + 1], a[4 * c + 2], a[4 * c + 3]);
  // a[4c + d] holds X[c + 4d]; transpose to natural order
  float2 t;
  t = a[1]; a[1] = a[4]; a[4] = t;
  t = a[2]; a[2] = a[8]; a[8] = t;
  t = a[3]; a[3] = a[12]; a[12] = t;
  t = a[6]; a[6] = a[9]; a[9] = t;
  t = a[7]; a[7] = a[13]; a[13] = t;
  t = a[11]; a[11] = a[14]; a[14] = t;
}

template <int R, bool INV> __device__ __forceinline__ void dftR(float2 (&a)[R]) {
  if constexpr (R == 2) dft2<INV>(a[0], a[1]);
  else if constexpr (R == 4) dft4<INV>(a[0], a[1], a[2], a[3]);
  else if constexpr (R == 8) dft8<INV>(a);
  else dft16<INV>(a);
}

// ------------------------------------------------------------------------------------
// One stage on the 16 register-resident points of thread t (t in [0, T)).
// Register v[u + U*m] <-> element index b*L + j + m*S, with beta = t + u*T, b = beta / S,
// j = beta % S, U = 16/R, S = L/R.
//   forward (DIF): butterfly then twiddle;  inverse (DIT): conj twiddle then butterfly.
// The twiddles of a stage depend only on the thread: they are fetched into registers by load_tw()
// ahead of the barrier of the preceding exchange (the data registers are dead at that point), so
// their L1/L2 latency overlaps the barrier wait instead of stalling the butterfly.
// ------------------------------------------------------------------------------------
template <int N, int S_IDX>
__device__ __forceinline__ void load_tw(float2 (&w)[kE], int t, const float2* __restrict__ tw_base) {
  constexpr int T = N / kE;
  constexpr int L = StageLen<N, S_IDX>::value;
  constexpr int R = Plan<N>::r[S_IDX];
  constexpr int S = L / R;
  constexpr int U = kE / R;
  if constexpr (S_IDX < Plan<N>::ns - 1) {
    const float2* tw = tw_base + TwOffset<N, S_IDX>::value;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = (t + u * T) & (S - 1);
#ifdef THZ_TW_POWERS
      // only the rows q = 1, 2, 4, 8 are fetched (a quarter of the table stays hot in L1), the other powers
      // are products of two fetched ones
#pragma unroll
      for (int q = 1; q < R; q <<= 1) w[u + U * q] = __ldg(&tw[(q - 1) * S + j]);
#pragma unroll
      for (int q = 3; q < R; ++q)
        if ((q & (q - 1)) != 0) {
          const int hi = (q >= 8) ? 8 : (q >= 4) ? 4 : 2;
          w[u + U * q] = cmul(w[u + U * hi], w[u + U * (q - hi)]);
        }
#else
#pragma unroll
      for (int q = 1; q < R; ++q) w[u + U * q] = __ldg(&tw[(q - 1) * S + j]);
#endif
    }
  }
}

template <int N, int S_IDX, bool INV>
__device__ __forceinline__ void stage_compute_tw(float2 (&v)[kE], const float2 (&w)[kE]) {
  constexpr int R = Plan<N>::r[S_IDX];
  constexpr int U = kE / R;
  constexpr bool kTw = (S_IDX < Plan<N>::ns - 1);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    float2 a[R];
#pragma unroll
    for (int m = 0; m < R; ++m) a[m] = v[u + U * m];
    if constexpr (INV && kTw) {
#pragma unroll
      for (int q = 1; q < R; ++q) a[q] = cmul_conj(a[q], w[u + U * q]);
    }
    dftR<R, INV>(a);
    if constexpr (!INV && kTw) {
#pragma unroll
      for (int q = 1; q < R; ++q) a[q] = cmul(a[q], w[u + U * q]);
    }
#pragma unroll
    for (int m = 0; m < R; ++m) v[u + U * m] = a[m];
  }
}

template <int N, int S_IDX, bool INV>
__device__ __forceinline__ void stage_compute(float2 (&v)[kE], int t, const float2* __restrict__ tw_base) {
  float2 w[kE];
  load_tw<N, S_IDX>(w, t, tw_base);
  stage_compute_tw<N, S_IDX, INV>(v, w);
}

// element index owned by register i = u + U*m of thread t at stage S_IDX
template <int N, int S_IDX>
__device__ __forceinline__ int stage_elem(int t, int i) {
  constexpr int T = N / kE;
  constexpr int L = StageLen<N, S_IDX>::value;
  constexpr int R = Plan<N>::r[S_IDX];
  constexpr int S = L / R;
  constexpr int U = kE / R;
  const int u = i % U, m = i / U;
  const int beta = t + u * T;
  return (beta / S) * L + (beta & (S - 1)) + m * S;
}

template <int N, int S_IDX>
__device__ __forceinline__ void stage_store(const float2 (&v)[kE], int t, float2* __restrict__ sm) {
#pragma unroll
  for (int i = 0; i < kE; ++i) sm[pad_idx(stage_elem<N, S_IDX>(t, i))] = v[i];
}
template <int N, int S_IDX>
__device__ __forceinline__ void stage_load(float2 (&v)[kE], int t, const float2* __restrict__ sm) {
#pragma unroll
  for (int i = 0; i < kE; ++i) v[i] = sm[pad_idx(stage_elem<N, S_IDX>(t, i))];
}

// natural-order bin k of the value at position p after the LAST forward stage:
// position p = sum_s q_s * S_s  <->  k = q_0 + r_0 q_1 + r_0 r_1 q_2 + ...
template <int N, int S_IDX = 0>
__device__ __forceinline__ int pos_to_bin(int p) {
  if constexpr (S_IDX == Plan<N>::ns) {
    return 0;
  } else {
    constexpr int L = StageLen<N, S_IDX>::value;
    constexpr int S = L / Plan<N>::r[S_IDX];
    constexpr int W = N / L;   // r_0 * ... * r_{s-1}
    const int q = p / S;
    return q * W + pos_to_bin<N, S_IDX + 1>(p - q * S);
  }
}

// Synchronisation of the exchanges.
//  * A group of T = N/16 threads only ever touches its own slice of the exchange buffer, so groups
//    that fit in one warp (N <= 512) synchronise with __syncwarp() and never stall the rest of the CTA.
//  * The exchange between stage s and s+1 moves data only inside aligned groups of S_s = L_s / R_s
//    consecutive threads when both stages run one radix-16 butterfly per thread (block b' = b R + q of
//    stage s+1 is written by threads [b S, (b+1) S) of stage s and read by threads inside the same
//    range).  When S_s <= 32 those groups live inside one warp: __syncwarp() is enough (N = 4096:
//    the exchange between the second and third stage; N = 8192: between the third and fourth).
//  * A stage stores to exactly the positions the same thread loaded at the end of the previous
//    exchange, so only the FIRST exchange of a transform needs a barrier in front of its stores
//    (other threads may still be reading the buffer from whatever used it before).
template <int N, int S_IDX> struct ExchangeScope {   // exchange between stage S_IDX and S_IDX + 1
  static constexpr int T = N / kE;
  static constexpr int R0 = Plan<N>::r[S_IDX], R1 = Plan<N>::r[S_IDX + 1];
  static constexpr int S = StageLen<N, S_IDX>::value / R0;
  static constexpr bool warp_local = (T <= 32) || (R0 == kE && R1 == kE && S <= 32);
};
template <bool WARP> __device__ __forceinline__ void scoped_sync() {
  if constexpr (WARP) __syncwarp();
  else __syncthreads();
}
template <int N> __device__ __forceinline__ void group_sync() { scoped_sync<(N / kE <= 32)>(); }

// Pull the slab of the NEXT work item into L2 while the current one is transformed: one 128-byte line per
// thread and step, no registers or shared memory held (the register-resident transforms cannot keep a second
// item in flight, and a shared-memory double buffer evicts the tables from L1).
__device__ __forceinline__ void prefetch_l2_slab(const float* base, int64_t floats) {
  const int64_t lines = (floats + 31) >> 5;
  for (int64_t l = threadIdx.x; l < lines; l += blockDim.x)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (l << 5)));
}

struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};

// Full forward transform: registers hold stage-0 layout (v[i] <-> element t + i*T) on entry,
// last-stage layout (digit-reversed positions) on exit.  `w` holds the twiddles of stage S_IDX
// (preloaded by the caller or by the previous exchange); `hook` runs between the stores and the
// barrier of the LAST exchange (callers prefetch the operands of their epilogue there).
template <int N, int S_IDX, typename Hook>
__device__ __forceinline__ void fft_forward_rec(float2 (&v)[kE], float2 (&w)[kE], int t, float2* sm,
                                                const float2* __restrict__ tw, Hook& hook) {
  stage_compute_tw<N, S_IDX, false>(v, w);
  if constexpr (S_IDX + 1 < Plan<N>::ns) {
    if constexpr (S_IDX == 0) group_sync<N>();   // previous users of sm are done
    stage_store<N, S_IDX>(v, t, sm);
    load_tw<N, S_IDX + 1>(w, t, tw);             // in flight across the barrier
    if constexpr (S_IDX + 2 == Plan<N>::ns) hook();
    scoped_sync<ExchangeScope<N, S_IDX>::warp_local>();
    stage_load<N, S_IDX + 1>(v, t, sm);
    fft_forward_rec<N, S_IDX + 1>(v, w, t, sm, tw, hook);
  }
}
template <int N, typename Hook>
__device__ __forceinline__ void fft_forward_hook(float2 (&v)[kE], int t, float2* sm, const float2* __restrict__ tw,
                                                 Hook& hook) {
  float2 w[kE];
  load_tw<N, 0>(w, t, tw);
  fft_forward_rec<N, 0>(v, w, t, sm, tw, hook);
}
template <int N>
__device__ __forceinline__ void fft_forward(float2 (&v)[kE], int t, float2* sm, const float2* __restrict__ tw) {
  NoHook h;
  fft_forward_hook<N>(v, t, sm, tw, h);
}

// Full inverse (unnormalised) transform: last-stage layout in, stage-0 layout out.
template <int N, int S_IDX, typename Hook>
__device__ __forceinline__ void fft_inverse_rec(float2 (&v)[kE], float2 (&w)[kE], int t, float2* sm,
                                                const float2* __restrict__ tw, Hook& hook) {
  stage_compute_tw<N, S_IDX, true>(v, w);
  if constexpr (S_IDX > 0) {
    if constexpr (S_IDX == Plan<N>::ns - 1) group_sync<N>();
    stage_store<N, S_IDX>(v, t, sm);
    load_tw<N, S_IDX - 1>(w, t, tw);
    if constexpr (S_IDX == 1) hook();
    scoped_sync<ExchangeScope<N, S_IDX - 1>::warp_local>();
    stage_load<N, S_IDX - 1>(v, t, sm);
    fft_inverse_rec<N, S_IDX - 1>(v, w, t, sm, tw, hook);
  }
}
template <int N, typename Hook>
__device__ __forceinline__ void fft_inverse_hook(float2 (&v)[kE], int t, float2* sm, const float2* __restrict__ tw,
                                                 Hook& hook) {
  float2 w[kE];   // the last stage has no twiddles
  fft_inverse_rec<N, Plan<N>::ns - 1>(v, w, t, sm, tw, hook);
}
template <int N>
__device__ __forceinline__ void fft_inverse(float2 (&v)[kE], int t, float2* sm, const float2* __restrict__ tw) {
  NoHook h;
  fft_inverse_hook<N>(v, t, sm, tw, h);
}

// ------------------------------------------------------------------------------------
// "is this trace entirely zero?" flags for a pair of traces packed as re / im.
// Packing two real traces into one complex transform leaks ~1e-7 of one trace into the other
// through rounding; the reference transforms traces separately, so an all-zero trace (a dead
// pixel) stays exactly zero there.  The kernels detect all-zero inputs at load time and write
// exact zeros for them.  Groups narrower than a warp resolve with one ballot; wider groups
// publish one word per warp to shared memory and read it back after the next barrier (the
// first exchange of the transform), double-buffered by item parity so that no extra barrier
// is needed.
// ------------------------------------------------------------------------------------
constexpr int kNzWords = 64;   // per group: 2 parities x 16 warps x 2 traces

template <int T>
__device__ __forceinline__ void nz_publish(bool nz0, bool nz1, int t, int g, int parity, unsigned* nzbuf, bool& z0,
                                           bool& z1) {
  const unsigned b0 = __ballot_sync(0xffffffffu, nz0), b1 = __ballot_sync(0xffffffffu, nz1);
  if constexpr (T < 32) {
    const unsigned lane0 = (threadIdx.x & 31u) / T * T;
    const unsigned mask = ((1u << T) - 1u) << lane0;
    z0 = (b0 & mask) == 0u;
    z1 = (b1 & mask) == 0u;
  } else {
    if ((t & 31) == 0) {
      unsigned* w = nzbuf + g * kNzWords + parity * 32 + (t >> 5) * 2;
      w[0] = b0;
      w[1] = b1;
    }
    z0 = z1 = false;   // resolved by nz_resolve after a barrier
  }
}

template <int T>
__device__ __forceinline__ void nz_resolve(int g, int parity, const unsigned* nzbuf, bool& z0, bool& z1) {
  if constexpr (T >= 32) {
    const unsigned* w = nzbuf + g * kNzWords + parity * 32;
    unsigned a0 = 0u, a1 = 0u;
#pragma unroll
    for (int i = 0; i < T / 32; ++i) {
      a0 |= w[2 * i];
      a1 |= w[2 * i + 1];
    }
    z0 = a0 == 0u;
    z1 = a1 == 0u;
  }
}

// ------------------------------------------------------------------------------------
// Bulk-copy staging of a CTA's input slab (TMA, cp.async.bulk global -> shared).
// The traces of one work item are contiguous in the [P][N] cube, so ONE elected thread moves the
// whole slab with a single asynchronous copy that completes on an mbarrier; the other threads
// spend no load instructions on it and the copy of item i+1 overlaps the transforms of item i.
// Two buffers: buffer (it & 1) holds item `it`; its refill for item it+2 is issued at the top of
// iteration it+1, after every thread has passed the barriers of iteration `it`.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct BulkStager {
  uint32_t mbar[2];
  __device__ __forceinline__ void init(uint64_t* bars) {
    mbar[0] = smem_addr(bars);
    mbar[1] = smem_addr(bars + 1);
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar[0]), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar[1]), "r"(1));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  // one thread: start the copy of `bytes` (multiple of 16, both addresses 16-byte aligned) into buffer b
  __device__ __forceinline__ void issue(int b, void* dst, const void* src, uint32_t bytes) const {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar[b]), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(mbar[b])
                 : "memory");
  }
  __device__ __forceinline__ void wait(int b, uint32_t phase) const {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
          : "=r"(ok)
          : "r"(mbar[b]), "r"(phase)
          : "memory");
    }
  }
};

}  // namespace thz
