// thz_trace.cu -- the per-pixel trace pass: window -> real FFT -> amplitude / unwrapped phase
// -> band-pass -> inverse FFT -> time gate -> intensity, as hand-written sm_100a kernels.
//
// Reference arithmetic being replaced (paths under the upstream repository):
//   src/math_tools.rs:330-398   fft()   (window, r2c, |s|, unwrap(arg s))
//   src/math_tools.rs:211-240   numpy_unwrap
//   src/math_tools.rs:546-567   ifft()  (c2r, / N)
//   src/filters/band_pass_fd.rs:155-212          (band multiplier on fft and amplitudes)
//   src/filters/band_pass_td_before_fft.rs:155-174, tilt_compensation.rs:188  (time multipliers)
//   src/data_thread.rs:1288-1307                 (intensity image)
//
// Layout: cube [P][N] f32, trace-major.  Two consecutive traces (2q, 2q+1) are packed as the
// real / imaginary part of one complex sequence and transformed by a group of T = N/16
// threads (thz_fft.cuh).  A CTA of 256 threads (512 for N = 8192) holds G = 256/T groups.
// Every trace crosses HBM once per kernel: coalesced 128-byte-per-warp loads straight into
// registers, streaming (evict-first) cache policy, multiplier vectors served from L1/L2.
#include "thz_trace_dev.cuh"

#include <math.h>

namespace thz {

// Same from the staged slab in shared memory (row 2g = trace p0, row 2g+1 = trace p0+1).
template <int N>
__device__ __forceinline__ void load_pair_staged(float2 (&v)[kE], const TraceArgs& a, const float* slab, int t, int g,
                                                 bool act0, bool act1, bool& nz0, bool& nz1) {
  constexpr int T = Geo<N>::T;
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

// ------------------------------------------------------------------------------------
// fused chain: one read and one write of the cube
// ------------------------------------------------------------------------------------
template <int N, bool STAGED>
__global__ void __launch_bounds__(Geo<N>::NT, Geo<N>::kMinBlocks) k_trace_fused(const TraceArgs a) {
  using GEO = Geo<N>;
  constexpr int T = GEO::T, G = GEO::G;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  float* scr = reinterpret_cast<float*>(smem + (size_t)G * padded_len(N));
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  float2* sm = smem + (size_t)g * padded_len(N);
  const int64_t npairs = (a.P + 1) >> 1;
  const int64_t nitems = (npairs + G - 1) / G;
  const bool use_post = a.m_post != nullptr;
  constexpr int LAST = Plan<N>::ns - 1;
  constexpr int RL = Plan<N>::r[LAST];
  constexpr int UL = kE / RL;
  unsigned* nzbuf = reinterpret_cast<unsigned*>(scr + 32 * G);
  int parity = 0;
  // input slabs arrive through bulk copies (thz_fft.cuh, BulkStager)
  float* slab[2] = {reinterpret_cast<float*>(smem_raw + GEO::stage_off),
                    reinterpret_cast<float*>(smem_raw + GEO::stage_off) + GEO::kSlabFloats};
  BulkStager stager;
  if constexpr (STAGED)
    stager.init(reinterpret_cast<uint64_t*>(smem_raw + GEO::stage_off + 2 * (size_t)GEO::kSlabFloats * sizeof(float)));
  auto slab_bytes = [&](int64_t it) -> uint32_t {
    const int64_t first = it * G * 2;
    int64_t cnt = a.P - first;
    if (cnt > 2 * G) cnt = 2 * G;
    return (uint32_t)(cnt * N * sizeof(float));
  };
  if (STAGED && threadIdx.x == 0 && (int64_t)blockIdx.x < nitems)
    stager.issue(0, slab[0], a.in + (int64_t)blockIdx.x * G * 2 * N, slab_bytes(blockIdx.x));
  uint32_t it_count = 0;

  // unstaged: the pair of an item is fetched into registers while the previous item is stored (loads in flight
  // across the stores and the intensity reduction), so that their L2 latency does not open every item
  float2 vn[STAGED ? 1 : kE];
  if constexpr (!STAGED) {
    if ((int64_t)blockIdx.x < nitems) {
      const int64_t q0 = ((int64_t)blockIdx.x * G + g) * 2;
      load_pair_raw<N>(vn, a.in, t, q0 < a.P, q0 + 1 < a.P, q0);
    }
  }
  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x, parity ^= 1, ++it_count) {
    const int64_t pair = item * G + g;
    const int64_t p0 = pair * 2;
    const bool act0 = p0 < a.P, act1 = p0 + 1 < a.P;
    const int buf = it_count & 1;
    const int64_t next = item + gridDim.x;
    // the refill below overwrites the slab every group read one iteration ago; groups that fit in a
    // warp only synchronise with __syncwarp() inside the transforms, so order them here
    float2 v[kE];
    bool nz0, nz1, z0, z1;
    if constexpr (STAGED) {
      if constexpr (T <= 32) __syncthreads();
      if (threadIdx.x == 0 && next < nitems)
        stager.issue(buf ^ 1, slab[buf ^ 1], a.in + next * G * 2 * N, slab_bytes(next));
      stager.wait(buf, (it_count >> 1) & 1);
      load_pair_staged<N>(v, a, slab[buf], t, g, act0, act1, nz0, nz1);
    } else {
      (void)buf;
      if (next < nitems) {
        int64_t cnt = a.P - next * G * 2;
        if (cnt > 2 * G) cnt = 2 * G;
        prefetch_l2_slab(a.in + next * G * 2 * N, cnt * N);
      }
#pragma unroll
      for (int i = 0; i < kE; ++i) v[i] = vn[i];
      finish_pair<N>(v, a, t, nz0, nz1);
    }
    nz_publish<T>(nz0, nz1, t, g, parity, nzbuf, z0, z1);
    // band-pass in digit-reversed order: register (u, m) <-> position (t + u*T)*RL + m, hq is stored
    // [m][beta] so that a warp reads consecutive floats; fetched while the last exchange is in flight
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
    fft_inverse<N>(v, t, sm, a.tw);
    nz_resolve<T>(g, parity, nzbuf, z0, z1);
    if constexpr (!STAGED) {
      {   // (past the last item the predicates are false: no loads, and vn does not stay live around the loop)
        const int64_t q0 = (next * G + g) * 2;
        const bool more = next < nitems;
        load_pair_raw<N>(vn, a.in, t, more && q0 < a.P, more && q0 + 1 < a.P, q0);
      }
    }
    store_pair<N>(v, a, t, g, act0, act1, p0, use_post, scr, z0, z1);
  }
}

// ------------------------------------------------------------------------------------
// forward: spectra materialised (drop-in for math_tools::fft)
// ------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(Geo<N>::NT, Geo<N>::kMinBlocks) k_trace_forward(const TraceArgs a) {
  using GEO = Geo<N>;
  constexpr int T = GEO::T, G = GEO::G;
  constexpr int F = N / 2 + 1;
  constexpr int PF = pad_idx(N / 2) + 1;   // padded floats per trace for the phase buffer
  constexpr int H = T / 2;                 // threads per trace in the unwrap
  constexpr int NU = (N / 2) / T + 1;      // bins per thread: k = t + u*T, u < NU (last only t == 0)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  float* scr = reinterpret_cast<float*>(smem + (size_t)G * padded_len(N));
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  float2* sm = smem + (size_t)g * padded_len(N);
  float* phs = reinterpret_cast<float*>(sm);
  const int64_t npairs = (a.P + 1) >> 1;
  const int64_t nitems = (npairs + G - 1) / G;
  constexpr int LAST = Plan<N>::ns - 1;
  const bool want_phase = a.phase != nullptr;
  unsigned* nzbuf = reinterpret_cast<unsigned*>(scr + 32 * G);
  int parity = 0;

  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x, parity ^= 1) {
    const int64_t pair = item * G + g;
    const int64_t p0 = pair * 2;
    const bool act0 = p0 < a.P, act1 = p0 + 1 < a.P;
    float2 v[kE];
    bool nz0, nz1, z0, z1;
    load_pair<N>(v, a, t, act0, act1, p0, nz0, nz1);
    nz_publish<T>(nz0, nz1, t, g, parity, nzbuf, z0, z1);
    if (a.win != nullptr) {   // the reference leaves the windowed trace in `data`
      float* w0 = a.win + p0 * N + t;
      float* w1 = w0 + N;
#pragma unroll
      for (int i = 0; i < kE; ++i) {
        if (act0) st_stream(w0 + i * T, v[i].x);
        if (act1) st_stream(w1 + i * T, v[i].y);
      }
    }
    fft_forward<N>(v, t, sm, a.tw);
    // scatter to natural bin order
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kE; ++i) sm[pad_idx(pos_to_bin<N>(stage_elem<N, LAST>(t, i)))] = v[i];
    __syncthreads();
    nz_resolve<T>(g, parity, nzbuf, z0, z1);
    // split the packed spectrum: X1 = (Z[k] + conj Z[N-k]) / 2, X2 = (Z[k] - conj Z[N-k]) / 2i
    float ph0[NU], ph1[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int k = t + u * T;
      ph0[u] = 0.f;
      ph1[u] = 0.f;
      if (u < NU - 1 || t == 0) {
        const float2 za = sm[pad_idx(k)];
        const float2 zb = sm[pad_idx((N - k) & (N - 1))];
        float2 x0 = make_float2(0.5f * (za.x + zb.x), 0.5f * (za.y - zb.y));
        float2 x1 = make_float2(0.5f * (za.y + zb.y), 0.5f * (zb.x - za.x));
        if (z0) x0 = make_float2(0.f, 0.f);   // all-zero trace: exact zero spectrum (and phase 0)
        if (z1) x1 = make_float2(0.f, 0.f);
        if (a.fft != nullptr) {
          if (act0) __stcs(a.fft + p0 * F + k, x0);
          if (act1) __stcs(a.fft + (p0 + 1) * F + k, x1);
        }
        if (a.amp != nullptr) {
          float a0 = sqrtf(fmaf(x0.x, x0.x, x0.y * x0.y)), a1 = sqrtf(fmaf(x1.x, x1.x, x1.y * x1.y));
          if (a.ref_amp != nullptr) {   // uniform over the grid
            const float r = fmaxf(__ldg(a.ref_amp + k), 1e-12f);
            a0 = a0 / r;
            a1 = a1 / r;
          }
          if (act0) st_stream(a.amp + p0 * F + k, a0);
          if (act1) st_stream(a.amp + (p0 + 1) * F + k, a1);
        }
        if (want_phase) {
          ph0[u] = atan2f(x0.y, x0.x);
          ph1[u] = atan2f(x1.y, x1.x);
        }
      }
    }
    if (want_phase) {   // uniform over the CTA
      __syncthreads();  // all reads of Z done; reuse the buffer for raw phases
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int k = t + u * T;
        if (u < NU - 1 || t == 0) {
          phs[pad_idx(k)] = ph0[u];
          phs[PF + pad_idx(k)] = ph1[u];
        }
      }
      __syncthreads();
      // threshold unwrap (src/math_tools.rs:224-237): thread (r, tau) owns bins 16 tau + 1 .. 16 tau + 16
      const int r = t / H, tau = t % H;
      float* pr = phs + r * PF;
      const float kPi = 3.14159265358979323846f, kTwoPi = 2.0f * kPi;
      float loc[16];
      float prev = pr[pad_idx(16 * tau)];
      float run = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float val = pr[pad_idx(16 * tau + 1 + c)];
        float d = val - prev;
        if (d > kPi) d -= kTwoPi;
        else if (d < -kPi) d += kTwoPi;
        run += d;
        loc[c] = run;
        prev = val;
      }
      // exclusive scan of `run` over tau (H threads, aligned to lanes)
      constexpr int W = (H < 32) ? H : 32;
      float inc = run;
#pragma unroll
      for (int o = 1; o < W; o <<= 1) {
        const float n = __shfl_up_sync(0xffffffffu, inc, o, W);
        if ((tau & (W - 1)) >= o) inc += n;
      }
      float base = inc - run;
      if constexpr (H > 32) {
        constexpr int NW = H / 32;
        float* s = scr + g * 32 + r * 16;
        if ((tau & 31) == 31) s[tau >> 5] = inc;
        __syncthreads();
        const int wq = tau >> 5;
#pragma unroll
        for (int w = 0; w < NW; ++w)
          if (w < wq) base += s[w];
      } else {
        __syncthreads();
      }
      base += pr[0];
      // all neighbours have read their `prev` (barrier above) -> overwrite in place
#pragma unroll
      for (int c = 0; c < 16; ++c) pr[pad_idx(16 * tau + 1 + c)] = base + loc[c];
      __syncthreads();
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int k = t + u * T;
        if (u < NU - 1 || t == 0) {
          const float rp = (a.ref_phase != nullptr) ? __ldg(a.ref_phase + k) : 0.f;
          if (act0) st_stream(a.phase + p0 * F + k, phs[pad_idx(k)] - rp);
          if (act1) st_stream(a.phase + (p0 + 1) * F + k, phs[PF + pad_idx(k)] - rp);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// inverse: spectra in (drop-in for math_tools::ifft, optionally fused with the FD band-pass
// and the time gate after the inverse FFT)
// ------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(Geo<N>::NT, Geo<N>::kMinBlocks) k_trace_inverse(const TraceArgs a) {
  using GEO = Geo<N>;
  constexpr int T = GEO::T, G = GEO::G;
  constexpr int F = N / 2 + 1;
  constexpr int NU = (N / 2) / T + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* smem = reinterpret_cast<float2*>(smem_raw);
  float* scr = reinterpret_cast<float*>(smem + (size_t)G * padded_len(N));
  const int g = threadIdx.x / T, t = threadIdx.x % T;
  float2* sm = smem + (size_t)g * padded_len(N);
  const int64_t npairs = (a.P + 1) >> 1;
  const int64_t nitems = (npairs + G - 1) / G;
  constexpr int LAST = Plan<N>::ns - 1;
  const bool use_post = a.m_post != nullptr;
  const float inv_n = 1.0f / (float)N;
  unsigned* nzbuf = reinterpret_cast<unsigned*>(scr + 32 * G);
  int parity = 0;

  for (int64_t item = blockIdx.x; item < nitems; item += gridDim.x, parity ^= 1) {
    const int64_t pair = item * G + g;
    const int64_t p0 = pair * 2;
    const bool act0 = p0 < a.P, act1 = p0 + 1 < a.P;
    bool nz0 = false, nz1 = false, z0, z1;
    __syncthreads();   // previous iteration's readers of sm are done
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int k = t + u * T;
      if (u < NU - 1 || t == 0) {
        float2 x0 = act0 ? __ldcs(a.fft_in + p0 * F + k) : make_float2(0.f, 0.f);
        float2 x1 = act1 ? __ldcs(a.fft_in + (p0 + 1) * F + k) : make_float2(0.f, 0.f);
        const float s = (a.band != nullptr) ? __ldg(a.band + k) * inv_n : inv_n;
        x0.x *= s; x0.y *= s; x1.x *= s; x1.y *= s;
        nz0 |= (x0.x != 0.f) | (x0.y != 0.f);
        nz1 |= (x1.x != 0.f) | (x1.y != 0.f);
        if (k == 0 || k == N / 2) {   // c2r ignores the imaginary parts of DC and Nyquist
          sm[pad_idx(k)] = make_float2(x0.x, x1.x);
        } else {
          sm[pad_idx(k)] = make_float2(x0.x - x1.y, x0.y + x1.x);
          sm[pad_idx(N - k)] = make_float2(x0.x + x1.y, x1.x - x0.y);
        }
      }
    }
    nz_publish<T>(nz0, nz1, t, g, parity, nzbuf, z0, z1);
    __syncthreads();
    float2 v[kE];
#pragma unroll
    for (int i = 0; i < kE; ++i) v[i] = sm[pad_idx(pos_to_bin<N>(stage_elem<N, LAST>(t, i)))];
    fft_inverse<N>(v, t, sm, a.tw);
    nz_resolve<T>(g, parity, nzbuf, z0, z1);
    store_pair<N>(v, a, t, g, act0, act1, p0, use_post, scr, z0, z1);
  }
}

// fft *= band, amp *= band (src/filters/band_pass_fd.rs:155-212); one thread per bin pair
__global__ void k_band_apply(float2* __restrict__ fft, float* __restrict__ amp, const float* __restrict__ band,
                             int64_t total, int F) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const float b = __ldg(band + (int)(i % F));
    if (fft != nullptr) {
      float2 x = fft[i];
      x.x *= b;
      x.y *= b;
      fft[i] = x;
    }
    if (amp != nullptr) amp[i] *= b;
  }
}

// out[p][t] = in[p][t] * mult[t] (mult may be null = ones); optional img[p] = sum_t out^2.
// One warp per trace, 128-bit accesses.
__global__ void k_time_multiply(const float* __restrict__ in, const float* __restrict__ mult, int n,
                                float* __restrict__ out, int64_t P, float* __restrict__ img) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < P; p += nwarps) {
    float acc = 0.f;
    if ((n & 3) == 0) {
      const float4* src = reinterpret_cast<const float4*>(in + p * n);
      float4* dst = out ? reinterpret_cast<float4*>(out + p * n) : nullptr;
      for (int q = lane; q < n / 4; q += 32) {
        float4 v = __ldcs(src + q);
        if (mult) {
          const float4 m = __ldg(reinterpret_cast<const float4*>(mult) + q);
          v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
        }
        if (dst) __stcs(dst + q, v);
        acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
      }
    } else {   // trace length not a multiple of 4: rows are not 16-byte aligned
      for (int i = lane; i < n; i += 32) {
        float v = __ldcs(in + p * n + i);
        if (mult) v *= __ldg(mult + i);
        if (out) __stcs(out + p * n + i, v);
        acc = fmaf(v, v, acc);
      }
    }
    if (img) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) img[p] = acc;
    }
  }
}

// `scale_3d` (src/math_tools.rs:270-298): out[nx][ny][z] = (sum_{i<s} sum_{j<s} in[nx*s+i][ny*s+j][z]) / (s*s),
// accumulated sequentially in the reference's order (i outer, j inner) so that the result is bit-exact.
__global__ void k_scale_blocks(const float* __restrict__ in, int width, int height, int zlen, int s, int new_w,
                               int new_h, float* __restrict__ out) {
  const int64_t total = (int64_t)new_w * new_h * zlen;
  const float factor = (float)(s * s);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int z = (int)(idx % zlen);
    const int64_t pix = idx / zlen;
    const int ny = (int)(pix % new_h), nx = (int)(pix / new_h);
    float sum = 0.f;
    for (int i = 0; i < s; ++i)
      for (int j = 0; j < s; ++j) {
        const int ox = nx * s + i, oy = ny * s + j;
        if (ox < width && oy < height) sum += __ldg(in + ((int64_t)ox * height + oy) * zlen + z);
      }
    out[idx] = sum / factor;
  }
}

// load path of `open_scan_from_thz` (src/io.rs:578-596): x <- x - x[0] per trace, img = sum x^2
__global__ void k_bias_subtract(const float* __restrict__ in, int n, float* __restrict__ out, int64_t P,
                                float* __restrict__ img) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < P; p += nwarps) {
    const float b = __ldg(in + p * n);
    float acc = 0.f;
    for (int i = lane; i < n; i += 32) {
      const float v = __ldcs(in + p * n + i) - b;
      __stcs(out + p * n + i, v);
      acc = fmaf(v, v, acc);
    }
    if (img) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) img[p] = acc;
    }
  }
}

// `average_polygon_roi` inner loops (src/math_tools.rs:636-659): out[z] = (sum over the listed pixels, in list
// order, of data[pix][z]) / count -- one thread per z keeps the reference's sequential f32 summation order.
__global__ void k_roi_average(const float* __restrict__ data, const int64_t* __restrict__ pix, int npix, int zlen,
                              float* __restrict__ out) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= zlen) return;
  float acc = 0.f;
  for (int i = 0; i < npix; ++i) acc += __ldg(data + pix[i] * zlen + z);
  out[z] = npix > 0 ? acc / (float)npix : acc;
}

// TiltCompensation with a non-zero tilt (src/filters/tilt_compensation.rs:158-199): every trace is tapered,
// shifted by its pixel's integer offset inside a time axis extended by 2 * num_steps samples; the head is
// filled with the trace's first RAW value, the tail with zeros.  insert[p] is computed on the host with the
// reference's mixed f32 / f64 arithmetic.
__global__ void k_tilt_shift(const float* __restrict__ in, const float* __restrict__ taper, const int* __restrict__ insert,
                             int n, int n_ext, int64_t P, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < P; p += nwarps) {
    const int ins = insert[p];
    const float first = __ldg(in + p * n);
    const int end = min(ins + n, n_ext);
    for (int k = lane; k < n_ext; k += 32) {
      float v;
      if (k < ins) v = first;
      else if (k < end) v = __ldcs(in + p * n + (k - ins)) * __ldg(taper + (k - ins));
      else v = 0.f;
      __stcs(out + p * n_ext + k, v);
    }
  }
}

// Voxel envelope of the 3-D view (src/gui/threed_plot.rs:165-204): one warp per trace; e[i] = sum_k ((x[j]^2)^contrast)
// * g[k], j = i + k - radius inside the trace (taps accumulated in order); traces whose maximum is below the
// opacity threshold are zeroed, the others min/max normalised.  The trace lives in shared memory.
__global__ void k_voxel_envelope(const float* __restrict__ in, int n, int64_t P, const float* __restrict__ kernel,
                                 int radius, float contrast, float opacity_threshold, float* __restrict__ out) {
  extern __shared__ float vsm[];   // per warp: n powered samples + n envelope values
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float* pw = vsm + (size_t)wib * 2 * n;
  float* ev = pw + n;
  for (int64_t p = (int64_t)blockIdx.x * wpb + wib; p < P; p += (int64_t)gridDim.x * wpb) {
    for (int i = lane; i < n; i += 32) {
      const float v = __ldcs(in + p * n + i);
      pw[i] = powf(v * v, contrast);
    }
    __syncwarp();
    float mx = -INFINITY, mn = INFINITY;
    for (int i = lane; i < n; i += 32) {
      float acc = 0.f;
      for (int k = 0; k <= 2 * radius; ++k) {
        const int j = i + k - radius;
        if (j >= 0 && j < n) acc += pw[j] * __ldg(kernel + k);
      }
      ev[i] = acc;
      mx = fmaxf(mx, acc);
      mn = fminf(mn, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    const float rng = mx - mn;
    const bool ok = (mx >= opacity_threshold) && (fabsf(rng) > 1e-6f);
    for (int i = lane; i < n; i += 32) __stcs(out + p * n + i, ok ? (ev[i] - mn) / rng : 0.f);
    __syncwarp();
  }
}

// histogram of the top 16 bits (pass 0) or, among the values whose top bits equal `prefix`, of the low 16 bits
// (pass 1) of non-negative floats: two passes give the exact k-th largest value (radix select)
__global__ void k_radix_hist(const float* __restrict__ x, int64_t total, int pass, unsigned prefix,
                             unsigned long long* __restrict__ hist) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned b = __float_as_uint(x[i]);
    if (pass == 0) atomicAdd(hist + (b >> 16), 1ull);
    else if ((b >> 16) == prefix) atomicAdd(hist + (b & 0xffffu), 1ull);
  }
}

// partial column sums of x[rows][cols]: block b sums rows [b*rpb, (b+1)*rpb) sequentially
__global__ void k_column_sums(const float* __restrict__ x, int64_t rows, int cols, float* __restrict__ partials) {
  const int64_t rpb = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * rpb;
  const int64_t r1 = (r0 + rpb < rows) ? r0 + rpb : rows;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float acc = 0.f;
    int64_t r = r0;
    for (; r + 4 <= r1; r += 4) {
      const float v0 = __ldcs(x + r * cols + c), v1 = __ldcs(x + (r + 1) * cols + c);
      const float v2 = __ldcs(x + (r + 2) * cols + c), v3 = __ldcs(x + (r + 3) * cols + c);
      acc += v0; acc += v1; acc += v2; acc += v3;
    }
    for (; r < r1; ++r) acc += __ldcs(x + r * cols + c);
    partials[(int64_t)blockIdx.x * cols + c] = acc;
  }
}

// ------------------------------------------------------------------------------------
// synthetic scan generator
// ------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

__global__ void k_generate(float* __restrict__ cube, int width, int height, int n, int row0, int total_width,
                           uint64_t seed, float t0, float dt, float noise) {
  const int64_t nq = (int64_t)width * height * (n / 4);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int q_per = n / 4;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
    const int64_t p = q / q_per;
    const int i0 = (int)(q % q_per) * 4;
    const int x = (int)(p / height) + row0, y = (int)(p % height);
    // amplitude: bars / checker pattern at two pitches so that deconvolution has structure
    const int c1 = ((x >> 4) + (y >> 4)) & 1, c2 = ((x >> 2) ^ (y >> 3)) & 1;
    const float amp = 1.0f + 0.35f * (c1 ? 1.f : -1.f) + 0.12f * (c2 ? 1.f : -1.f);
    const float up = 0.5f + 0.5f * __sinf(0.013f * (float)x + 0.021f * (float)y);
    const float tp = 10.0f + 5.0f * up;   // pulse position relative to t0 (ps)
    const float tau = 0.3f, fc = 1.0f;
    const uint64_t gp = (uint64_t)x * (uint64_t)height + (uint64_t)y;
    const uint64_t h0 = mix64(seed ^ mix64(gp * 0x100000001B3ull + (uint64_t)(i0 >> 2)));
    const uint64_t h1 = mix64(h0);
    float nz[4];
    {
      const float u0 = ((float)(uint32_t)(h0 & 0xffffffffu) + 0.5f) * (1.0f / 4294967296.0f);
      const float u1 = ((float)(uint32_t)(h0 >> 32) + 0.5f) * (1.0f / 4294967296.0f);
      const float u2 = ((float)(uint32_t)(h1 & 0xffffffffu) + 0.5f) * (1.0f / 4294967296.0f);
      const float u3 = ((float)(uint32_t)(h1 >> 32) + 0.5f) * (1.0f / 4294967296.0f);
      const float ra = sqrtf(-2.0f * __logf(u0)), rb = sqrtf(-2.0f * __logf(u2));
      float s, c;
      __sincosf(6.283185307179586f * u1, &s, &c);
      nz[0] = ra * c; nz[1] = ra * s;
      __sincosf(6.283185307179586f * u3, &s, &c);
      nz[2] = rb * c; nz[3] = rb * s;
    }
    float4 o;
    float* op = &o.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float tt = (float)(i0 + j) * dt - tp;
      const float e = tt / tau;
      op[j] = amp * __expf(-e * e) * __cosf(6.283185307179586f * fc * tt) + noise * nz[j];
    }
    (void)t0;
    (void)total_width;
    reinterpret_cast<float4*>(cube)[q] = o;
  }
}

// ------------------------------------------------------------------------------------
// host side: tables and launchers
// ------------------------------------------------------------------------------------
bool supported_n(int n) {
  return n == 64 || n == 128 || n == 256 || n == 512 || n == 1024 || n == 2048 || n == 4096 || n == 8192;
}

template <int N> static void host_plan(int& ns, int (&r)[4]) {
  ns = Plan<N>::ns;
  for (int i = 0; i < 4; ++i) r[i] = Plan<N>::r[i];
}
static bool plan_of(int n, int& ns, int (&r)[4]) {
  switch (n) {
    case 64: host_plan<64>(ns, r); return true;
    case 128: host_plan<128>(ns, r); return true;
    case 256: host_plan<256>(ns, r); return true;
    case 512: host_plan<512>(ns, r); return true;
    case 1024: host_plan<1024>(ns, r); return true;
    case 2048: host_plan<2048>(ns, r); return true;
    case 4096: host_plan<4096>(ns, r); return true;
    case 8192: host_plan<8192>(ns, r); return true;
    default: return false;
  }
}

int build_twiddles(int n, std::vector<float2>& tw) {
  int ns, r[4];
  if (!plan_of(n, ns, r)) return THZ_EINVAL;
  tw.clear();
  int L = n;
  for (int s = 0; s + 1 < ns; ++s) {
    const int R = r[s], S = L / R;
    for (int q = 1; q < R; ++q)
      for (int j = 0; j < S; ++j) {
        const double ang = -2.0 * M_PI * (double)j * (double)q / (double)L;
        tw.push_back(make_float2((float)cos(ang), (float)sin(ang)));
      }
    L = S;
  }
  if (tw.empty()) tw.push_back(make_float2(1.f, 0.f));
  return THZ_OK;
}

static int host_pos_to_bin(int n, int ns, const int* r, int p) {
  int k = 0, w = 1, L = n;
  for (int s = 0; s < ns; ++s) {
    const int S = L / r[s];
    const int q = p / S;
    p -= q * S;
    k += q * w;
    w *= r[s];
    L = S;
  }
  return k;
}

// hq[m * (n / RL) + beta] = H[k(beta * RL + m)] / n with H the band multiplier mirrored to all n bins
int build_hq(int n, const float* band, std::vector<float>& hq) {
  int ns, r[4];
  if (!plan_of(n, ns, r)) return THZ_EINVAL;
  const int RL = r[ns - 1];
  hq.assign(n, 0.f);
  const float inv_n = 1.0f / (float)n;
  for (int beta = 0; beta < n / RL; ++beta)
    for (int m = 0; m < RL; ++m) {
      const int k = host_pos_to_bin(n, ns, r, beta * RL + m);
      const int kk = (k <= n / 2) ? k : n - k;
      const float h = band ? band[kk] : 1.0f;
      hq[(size_t)m * (n / RL) + beta] = h * inv_n;
    }
  return THZ_OK;
}

template <int N, typename K>
static int launch_geo(thz_ctx* c, cudaStream_t s, K kernel, const TraceArgs& a, bool staged = false,
                      int carveout_pct = -1) {
  using GEO = Geo<N>;
  const size_t smem = staged ? GEO::smem_bytes_staged : GEO::smem_bytes;
  const void* key = (const void*)kernel;
  auto it = c->occ.find(key);
  if (it == c->occ.end()) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaFuncSetAttribute(smem)");
    if (carveout_pct >= 0) {   // leave the rest of the unified L1 / shared memory to the twiddle and multiplier tables
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_pct);
      if (e != cudaSuccess) return cuda_fail(c, e, "cudaFuncSetAttribute(carveout)");
    }
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, GEO::NT, smem);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (nb < 1) return set_err(c, THZ_ECUDA, "trace kernel does not fit on an SM");
    it = c->occ.emplace(key, nb).first;
  }
  const int64_t npairs = (a.P + 1) / 2;
  const int64_t nitems = (npairs + GEO::G - 1) / GEO::G;
  if (nitems <= 0) return THZ_OK;
  int64_t grid = (int64_t)c->sm_count * it->second;   // persistent: one wave of resident CTAs
  if (grid > nitems) grid = nitems;
  kernel<<<(unsigned)grid, GEO::NT, smem, s>>>(a);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "trace kernel launch");
  return THZ_OK;
}

#define THZ_DISPATCH_N(n, FN, ...)                 \
  switch (n) {                                     \
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

template <int N> static int do_fused(thz_ctx* c, cudaStream_t s, const TraceArgs& a) {
  if (c->unstaged_fir) return launch_geo<N>(c, s, k_trace_fused<N, false>, a, false, 40);
  return launch_geo<N>(c, s, k_trace_fused<N, true>, a, true);
}
template <int N> static int do_forward(thz_ctx* c, cudaStream_t s, const TraceArgs& a) {
  return launch_geo<N>(c, s, k_trace_forward<N>, a);
}
template <int N> static int do_inverse(thz_ctx* c, cudaStream_t s, const TraceArgs& a) {
  return launch_geo<N>(c, s, k_trace_inverse<N>, a);
}

static int base_args(thz_ctx* c, TraceArgs& a, int64_t P) {
  if (c->plan.n == 0) return set_err(c, THZ_ESTATE, "thz_plan_trace has not been called");
  if (P < 0) return set_err(c, THZ_EINVAL, "negative trace count");
  const FftTables* tb = nullptr;
  int rc = get_tables(c, c->plan.n, &tb);
  if (rc != THZ_OK) return rc;
  a = TraceArgs{};
  a.tw = tb->d_tw;
  a.P = P;
  return THZ_OK;
}

int trace_fused_args(thz_ctx* c, TraceArgs& a, const float* d_in, float* d_out, float* d_img, int64_t P) {
  int rc = base_args(c, a, P);
  if (rc != THZ_OK) return rc;
  if (P == 0) return THZ_OK;
  if (!d_in || !d_out) return set_err(c, THZ_EINVAL, "null cube pointer");
  if ((reinterpret_cast<uintptr_t>(d_in) & 15u) != 0) return set_err(c, THZ_EINVAL, "cube must be 16-byte aligned");
  a.in = d_in;
  a.out = d_out;
  a.img = d_img;
  a.m_pre = c->plan.has_pre ? c->plan.d_m_pre : nullptr;
  // a gate of ones (post_mode 0) multiplies by exactly 1.0f: skipped
  a.m_post = (c->plan.has_post && c->plan.post_mode != 0) ? c->plan.d_m_post : nullptr;
  a.pre_ends = c->plan.pre_ends_only ? 1 : 0;
  a.post_ends = c->plan.post_ends_only ? 1 : 0;
  a.hq = c->plan.d_hq;
  return THZ_OK;
}

int launch_trace_fused(thz_ctx* c, cudaStream_t s, const float* d_in, float* d_out, float* d_img, int64_t P) {
  if (c->plan.n != 0 && c->plan.blue_m != 0) {
    if (P == 0) return THZ_OK;
    if (!d_in || !d_out) return set_err(c, THZ_EINVAL, "null cube pointer");
    return launch_blue_fused(c, s, d_in, d_out, d_img, P);
  }
  TraceArgs a;
  int rc = trace_fused_args(c, a, d_in, d_out, d_img, P);
  if (rc != THZ_OK) return rc;
  if (P == 0) return THZ_OK;
  THZ_DISPATCH_N(c->plan.n, do_fused, c, s, a);
}

// amp[p][k] /= max(ref_amp[k], 1e-12), phase[p][k] -= ref_phase[k] on [P][F] arrays (chirp-z plans, whose forward
// kernel has no normalising epilogue)
__global__ void k_reference_normalise(float* __restrict__ amp, float* __restrict__ phase, const float* __restrict__ ref_amp,
                                      const float* __restrict__ ref_phase, int64_t total, int F) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int k = (int)(i % F);
    if (amp) amp[i] = amp[i] / fmaxf(__ldg(ref_amp + k), 1e-12f);
    if (phase) phase[i] = phase[i] - __ldg(ref_phase + k);
  }
}

// out[p] = a[p][bin] for a [P][F] array: one map slice of a spectral cube
__global__ void k_spectral_slice(const float* __restrict__ a, int F, int bin, float* __restrict__ out, int64_t P) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += stride) out[p] = __ldg(a + p * F + bin);
}

int launch_spectral_slice(thz_ctx* c, cudaStream_t s, const float* d_a, int F, int bin, float* d_out, int64_t P) {
  if (P == 0) return THZ_OK;
  const int blocks = (int)std::min<int64_t>((P + 255) / 256, (int64_t)c->sm_count * 8);
  k_spectral_slice<<<blocks, 256, 0, s>>>(d_a, F, bin, d_out, P);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "k_spectral_slice launch");
  return THZ_OK;
}

int launch_trace_forward(thz_ctx* c, cudaStream_t s, const float* d_in, float* d_win, float2* d_fft, float* d_amp,
                         float* d_phase, int64_t P, bool normalise) {
  if (normalise && (!c->plan.d_ref_amp || !c->plan.d_ref_phase || c->plan.ref_f != c->plan.n / 2 + 1))
    return set_err(c, THZ_ESTATE, "thz_plan_reference has not been called for this trace length");
  if (c->plan.n != 0 && c->plan.blue_m != 0) {
    if (P == 0) return THZ_OK;
    if (!d_in) return set_err(c, THZ_EINVAL, "null cube pointer");
    int rc = launch_blue_forward(c, s, d_in, d_win, d_fft, d_amp, d_phase, P);
    if (rc == THZ_OK && normalise && (d_amp || d_phase)) {
      const int F = c->plan.n / 2 + 1;
      const int64_t total = P * F;
      const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)c->sm_count * 8);
      k_reference_normalise<<<blocks, 256, 0, s>>>(d_amp, d_phase, c->plan.d_ref_amp, c->plan.d_ref_phase, total, F);
      c->launches++;
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return cuda_fail(c, e, "k_reference_normalise launch");
    }
    return rc;
  }
  TraceArgs a;
  int rc = base_args(c, a, P);
  if (rc != THZ_OK) return rc;
  if (P == 0) return THZ_OK;
  if (!d_in) return set_err(c, THZ_EINVAL, "null cube pointer");
  a.in = d_in;
  a.win = d_win;
  a.fft = d_fft;
  a.amp = d_amp;
  a.phase = d_phase;
  a.m_pre = c->plan.has_pre ? c->plan.d_m_pre : nullptr;
  a.pre_ends = c->plan.pre_ends_only ? 1 : 0;
  a.ref_amp = normalise ? c->plan.d_ref_amp : nullptr;
  a.ref_phase = normalise ? c->plan.d_ref_phase : nullptr;
  THZ_DISPATCH_N(c->plan.n, do_forward, c, s, a);
}

int launch_trace_inverse(thz_ctx* c, cudaStream_t s, const float2* d_fft, bool use_band, bool use_post, float* d_out,
                         float* d_img, int64_t P) {
  if (c->plan.n != 0 && c->plan.blue_m != 0) {
    if (P == 0) return THZ_OK;
    if (!d_fft || !d_out) return set_err(c, THZ_EINVAL, "null pointer");
    return launch_blue_inverse(c, s, d_fft, use_band, use_post, d_out, d_img, P);
  }
  TraceArgs a;
  int rc = base_args(c, a, P);
  if (rc != THZ_OK) return rc;
  if (P == 0) return THZ_OK;
  if (!d_fft || !d_out) return set_err(c, THZ_EINVAL, "null pointer");
  a.fft_in = d_fft;
  a.out = d_out;
  a.img = d_img;
  a.band = (use_band && c->plan.has_band) ? c->plan.d_band : nullptr;
  a.m_post = (use_post && c->plan.has_post) ? c->plan.d_m_post : nullptr;
  a.post_ends = c->plan.post_ends_only ? 1 : 0;
  THZ_DISPATCH_N(c->plan.n, do_inverse, c, s, a);
}

int launch_band_apply(thz_ctx* c, cudaStream_t s, float2* d_fft, float* d_amp, int64_t P) {
  if (c->plan.n == 0) return set_err(c, THZ_ESTATE, "thz_plan_trace has not been called");
  if (!c->plan.has_band || P == 0) return THZ_OK;   // band of ones
  const int F = c->plan.n / 2 + 1;
  const int64_t total = P * F;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)c->sm_count * 16;
  if (blocks > cap) blocks = cap;
  k_band_apply<<<(unsigned)blocks, 256, 0, s>>>(d_fft, d_amp, c->plan.d_band, total, F);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "k_band_apply launch");
  return THZ_OK;
}

int launch_time_multiply(thz_ctx* c, cudaStream_t s, const float* d_in, const float* d_mult, int n, float* d_out,
                         int64_t P, float* d_img) {
  if (P == 0) return THZ_OK;
  int64_t blocks = (P * 32 + 255) / 256;
  const int64_t cap = (int64_t)c->sm_count * 16;
  if (blocks > cap) blocks = cap;
  k_time_multiply<<<(unsigned)blocks, 256, 0, s>>>(d_in, d_mult, n, d_out, P, d_img);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "k_time_multiply launch");
  return THZ_OK;
}

int launch_scale_blocks(thz_ctx* c, cudaStream_t s, const float* d_in, int width, int height, int zlen, int sf,
                        float* d_out) {
  const int new_w = width / sf, new_h = height / sf;
  const int64_t total = (int64_t)new_w * new_h * zlen;
  if (total == 0) return THZ_OK;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)c->sm_count * 16;
  if (blocks > cap) blocks = cap;
  k_scale_blocks<<<(unsigned)blocks, 256, 0, s>>>(d_in, width, height, zlen, sf, new_w, new_h, d_out);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "k_scale_blocks launch");
  return THZ_OK;
}

int launch_bias_subtract(thz_ctx* c, cudaStream_t s, const float* d_in, int n, float* d_out, int64_t P, float* d_img) {
  if (P == 0) return THZ_OK;
  int64_t blocks = (P * 32 + 255) / 256;
  const int64_t cap = (int64_t)c->sm_count * 16;
  if (blocks > cap) blocks = cap;
  k_bias_subtract<<<(unsigned)blocks, 256, 0, s>>>(d_in, n, d_out, P, d_img);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "k_bias_subtract launch");
  return THZ_OK;
}

int launch_roi_average(thz_ctx* c, cudaStream_t s, const float* d_data, const int64_t* d_pix, int npix, int zlen,
                       float* d_out) {
  if (zlen <= 0) return THZ_OK;
  k_roi_average<<<(zlen + 127) / 128, 128, 0, s>>>(d_data, d_pix, npix, zlen, d_out);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "k_roi_average launch");
  return THZ_OK;
}

int launch_tilt_shift(thz_ctx* c, cudaStream_t s, const float* d_in, const float* d_taper, const int* d_insert, int n,
                      int n_ext, int64_t P, float* d_out) {
  if (P == 0) return THZ_OK;
  int64_t blocks = (P * 32 + 255) / 256;
  const int64_t cap = (int64_t)c->sm_count * 16;
  if (blocks > cap) blocks = cap;
  k_tilt_shift<<<(unsigned)blocks, 256, 0, s>>>(d_in, d_taper, d_insert, n, n_ext, P, d_out);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "k_tilt_shift launch");
  return THZ_OK;
}

int launch_voxel_envelope(thz_ctx* c, cudaStream_t s, const float* d_in, int n, int64_t P, const float* d_kernel,
                          int radius, float contrast, float thr, float* d_out) {
  if (P == 0) return THZ_OK;
  const int wpb = 4;
  const size_t smem = (size_t)wpb * 2 * n * sizeof(float);
  static size_t have = 0;
  if (smem > have) {
    cudaError_t e = cudaFuncSetAttribute(k_voxel_envelope, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaFuncSetAttribute(voxel)");
    have = smem;
  }
  int64_t blocks = (P + wpb - 1) / wpb;
  const int64_t cap = (int64_t)c->sm_count * 8;
  if (blocks > cap) blocks = cap;
  k_voxel_envelope<<<(unsigned)blocks, wpb * 32, smem, s>>>(d_in, n, P, d_kernel, radius, contrast, thr, d_out);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "k_voxel_envelope launch");
  return THZ_OK;
}

int launch_radix_hist(thz_ctx* c, cudaStream_t s, const float* d_x, int64_t total, int pass, unsigned prefix,
                      unsigned long long* d_hist) {
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)c->sm_count * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  k_radix_hist<<<(unsigned)blocks, 256, 0, s>>>(d_x, total, pass, prefix, d_hist);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "k_radix_hist launch");
  return THZ_OK;
}

int launch_column_sums(thz_ctx* c, cudaStream_t s, const float* d_x, int64_t rows, int cols, float* d_partials,
                       int nblocks) {
  k_column_sums<<<nblocks, 256, 0, s>>>(d_x, rows, cols, d_partials);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "k_column_sums launch");
  return THZ_OK;
}

int launch_generate(thz_ctx* c, cudaStream_t s, float* d_cube, int width, int height, int n, int row0,
                    int total_width, uint64_t seed, float t0, float dt, float noise) {
  if (n % 4 != 0) return set_err(c, THZ_EINVAL, "n must be a multiple of 4");
  const int64_t nq = (int64_t)width * height * (n / 4);
  if (nq == 0) return THZ_OK;
  int64_t blocks = (nq + 255) / 256;
  const int64_t cap = (int64_t)c->sm_count * 32;
  if (blocks > cap) blocks = cap;
  k_generate<<<(unsigned)blocks, 256, 0, s>>>(d_cube, width, height, n, row0, total_width, seed, t0, dt, noise);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "k_generate launch");
  return THZ_OK;
}

}  // namespace thz
