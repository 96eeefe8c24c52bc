// thz_rl.cu -- Richardson-Lucy on sm_100a: TMA-staged separable / dense 2-D filtering, the streaming strip
// kernel, and its row-slab form with halo rows pushed between GPUs over NVLink.
//
// Reference (paths under the upstream repository):
//   src/filters/deconvolution.rs:620-712           richardson_lucy
//   src/filters/deconvolution.rs:432-545           direct_convolve2d / FFT convolve2d
//   src/filters/deconvolution.rs:975, 990-993      clamp, gain = sqrt(u / d)
#include <cuda.h>
#include <cudaTypedefs.h>

#include "thz_internal.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <initializer_list>
#include <type_traits>
#include <vector>

namespace thz {
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

// ---- row-slab form (several GPUs, one slab of the padded image each) -------------------------------------
// The image a kernel filters lives in a buffer of halo + own + halo rows: the own rows of this rank and the
// `halo` = kx/2 boundary rows of the two neighbours, which the NEIGHBOURS' kernels store there directly over
// NVLink (peer pointers; cudaIpc-mapped when the ranks are processes) while they write their own rows.  The
// output rows are cut into explicit segments: segment 0 = the top `halo` own rows, the last segment = the
// bottom `halo` own rows, interior segments in between.  Exactly the CTAs of the two boundary segments
//   * read halo rows, so only they wait (one thread, ld.acquire.sys) for the neighbour's flag of the image
//     version this kernel consumes -- the interior segments start at once and hide the NVLink latency;
//   * produce the rows the neighbours need: they push them, fence, and the last of them to finish (a local
//     counter) publishes the version number in the neighbour's flag (st.release.sys).
// Readers of a halo and producers of the next version of the same rows are the same CTAs, so a neighbour that
// has seen version v of our rows may overwrite the halo it pushed for version v - 1: no extra handshake.
constexpr int kMaxSlabSegs = 40;

struct NoSlab {};

struct SlabArgs {
  int row_off;                       // first own row inside the mapped buffer (= halo)
  int halo, own;
  int seg_start[kMaxSlabSegs + 1];   // own-row ranges of the segments, seg_start[nseg] = own
  float* up_out;                     // rank above: row 0 of ITS bottom halo of the image this kernel writes (null: none)
  float* down_out;                   // rank below: row 0 of ITS top halo
  const unsigned long long* wait_up;     // local flags: version of the image this kernel READS that the neighbour
  const unsigned long long* wait_down;   // has pushed into our halos
  unsigned long long wait_val;
  unsigned long long* cnt_up;        // local counters of boundary CTAs that finished pushing
  unsigned long long* cnt_down;
  unsigned long long cnt_target;     // counter value that marks the last CTA of this launch
  unsigned long long* sig_up;        // flags in the neighbours' memory for the image this kernel WRITES
  unsigned long long* sig_down;
  unsigned long long sig_val;
  int* err;                          // set when a wait times out (a peer died): the host reports THZ_ECUDA
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// one thread: wait until *flag >= want; gives up after ~4 s so that a dead peer cannot hang the GPU
__device__ __forceinline__ void slab_wait(const unsigned long long* flag, unsigned long long want, int* err) {
  if (ld_acquire_sys(flag) >= want) return;
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(flag) < want) {
    __nanosleep(64);
    if (global_ns() - t0 > 4000000000ull) {
      if (err) *err = 1;
      break;
    }
  }
}
// all threads of a CTA that pushed rows to a neighbour: make the stores visible, count the CTA, and let the
// last one publish the version
__device__ __forceinline__ void slab_signal(unsigned long long* cnt, unsigned long long target,
                                            unsigned long long* sig, unsigned long long val) {
  // the CTA barrier orders every thread's peer stores before thread 0's system-scope fence, which is cumulative:
  // one fence per CTA instead of one per thread (256 fences behind NVLink stores cost 6 us, measured)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(cnt, 1ull) + 1ull == target) {
      __threadfence_system();
      st_release_sys(sig, val);
    }
  }
}

// launch slot -> segment: slot 0 = first segment (top boundary), slot 1 = last segment (bottom boundary), then the
// interior segments in order
__device__ __forceinline__ int slab_segment_of_slot(int slot, int nseg) {
  return slot == 0 ? 0 : (slot == 1 ? nseg - 1 : slot - 1);
}

// what one CTA needs of the slab bookkeeping, in registers (the segment table stays where it is)
struct SlabCore {
  int row_off = 0, halo = 0, own = 0;
  bool top = false, bottom = false;          // this CTA belongs to the first / last segment
  float* up_out = nullptr;
  float* down_out = nullptr;
  const unsigned long long* wait_up = nullptr;
  const unsigned long long* wait_down = nullptr;
  unsigned long long wait_val = 0;
  unsigned long long* cnt_up = nullptr;
  unsigned long long* cnt_down = nullptr;
  unsigned long long cnt_target = 0;
  unsigned long long* sig_up = nullptr;
  unsigned long long* sig_down = nullptr;
  unsigned long long sig_val = 0;
  int* err = nullptr;
  unsigned long long* dbg = nullptr;   // THZ_SLAB_TRACE: {first start, longest halo wait, last end, last boundary end,
                                       //  strip-0 top CTA: t0 t1 t2 t3 t4 -, strip-0 first interior CTA: t0 t1 t2 t3 t4 -} (ns)
  int dbg_slot = -1;
};

// The strip march itself: CTA (strip bx, output rows [seg_row0, seg_row0 + rows_out)) of one filtering.
template <int MODE, int SW, bool SLAB>
__device__ __forceinline__ void rl_stream_body(const CUtensorMap* tmap, const StreamArgs& a, const SlabCore& sl, int bx,
                                               int seg_row0, int rows_out) {
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
  const int col0 = bx * kSW + a.col_shift;
  const int row_off = SLAB ? sl.row_off : 0;
  const int n_chunks = (rows_out + a.WU + kCR - 1) / kCR;
  const int gy0 = row_off + seg_row0 - a.gy_off, gx0 = col0 - a.gx_off;
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
    if constexpr (SLAB) {
      // boundary segments read the neighbours' rows: wait for the version this kernel consumes (the other
      // threads block on the first chunk's mbarrier); the TMA reads below are ordered after the acquire
      const unsigned long long tw0 = sl.dbg ? global_ns() : 0ull;
      if (sl.top && sl.wait_up) slab_wait(sl.wait_up, sl.wait_val, sl.err);
      if (sl.bottom && sl.wait_down) slab_wait(sl.wait_down, sl.wait_val, sl.err);
      asm volatile("fence.proxy.async.global;" ::: "memory");
      if (sl.dbg) {
        const unsigned long long tw1 = global_ns();
        atomicMin(sl.dbg, tw0);
        if (sl.top || sl.bottom) atomicMax(sl.dbg + 1, tw1 - tw0);
        if (bx == 0 && (sl.dbg_slot == 0 || sl.dbg_slot == 2)) {
          unsigned long long* f = sl.dbg + (sl.dbg_slot == 0 ? 4 : 10);
          f[0] = tw0;
          f[1] = tw1;
        }
      }
    }
    for (int j = 0; j < 2 && j < n_chunks; ++j)
      if (live(j)) {
        mbar_expect_tx(mbar0 + 8 * j, tile_bytes);
        tma_load_2d(smem_u32(tile0 + j * tile_floats), tmap, gx0, gy0 + kCR * j, mbar0 + 8 * j);
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
    const size_t o0 = (size_t)(row_off + seg_row0 + i0) * a.pitch + gc;
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
      // rows past rows_out + WU feed no output of this segment (the last chunk of a short segment is mostly such rows)
      const bool need = j * kCR + r < rows_out + a.WU;
      if (lv) {
        if (need) {
          const uint32_t mb = mbar0 + 8 * buf;
          while (!mbar_try_wait(mb, buf ? ph1 : ph0)) {
          }
          LinearLoader ld{tile0 + buf * tile_floats + r * a.bc + cg * 16};
          run_taps(acc, ld, wys, a.KW);
        }
        if (buf) ph1 ^= 1; else ph0 ^= 1;
        if constexpr (SLAB) {
          if (sl.dbg && j == 0 && tid == 0 && bx == 0 && (sl.dbg_slot == 0 || sl.dbg_slot == 2))
            sl.dbg[(sl.dbg_slot == 0 ? 4 : 10) + 2] = global_ns();     // first tile has landed, column pass done
        }
      }
      if (need) {
        int pos = (j * kCR + r) % a.Rg;
        float* dst = ring + (cg * 16) * a.RS + pos;
#pragma unroll
        for (int q = 0; q < 16; ++q) dst[q * a.RS] = acc[q];
      }
    }
    __syncthreads();   // ring rows of chunk j are visible; tile[buf] is free
    if (tid == 0 && j + 2 < n_chunks && live(j + 2)) {
      mbar_expect_tx(mbar0 + 8 * buf, tile_bytes);
      tma_load_2d(smem_u32(tile0 + buf * tile_floats), tmap, gx0, gy0 + kCR * (j + 2), mbar0 + 8 * buf);
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
            float v;
            if constexpr (MODE == 0) v = acc[q];
            else if constexpr (MODE == 1) v = ep[q] / (acc[q] + a.eps);
            else v = ep[q] * acc[q];
            a.out[o] = v;
            if constexpr (SLAB) {
              // own rows [0, halo) are the bottom halo of the rank above, own rows [own - halo, own) the top
              // halo of the rank below: stored there as well (128-byte coalesced NVLink writes per warp)
              const int orow = seg_row0 + i0 + q;
              if (sl.up_out && orow < sl.halo) sl.up_out[(size_t)orow * a.pitch + gc] = v;
              if (sl.down_out && orow >= sl.own - sl.halo)
                sl.down_out[(size_t)(orow - (sl.own - sl.halo)) * a.pitch + gc] = v;
            }
          }
        }
      }
    }
  }
  if constexpr (SLAB) {
    if (sl.dbg && tid == 0 && bx == 0 && (sl.dbg_slot == 0 || sl.dbg_slot == 2))
      sl.dbg[(sl.dbg_slot == 0 ? 4 : 10) + 3] = global_ns();           // all chunks done
    if (sl.top && sl.sig_up) slab_signal(sl.cnt_up, sl.cnt_target, sl.sig_up, sl.sig_val);
    if (sl.bottom && sl.sig_down) slab_signal(sl.cnt_down, sl.cnt_target, sl.sig_down, sl.sig_val);
    if (sl.dbg && tid == 0) {
      const unsigned long long te = global_ns();
      atomicMax(sl.dbg + 2, te);
      if (sl.top || sl.bottom) atomicMax(sl.dbg + 3, te);   // last boundary CTA done (signal sent)
      if (bx == 0 && (sl.dbg_slot == 0 || sl.dbg_slot == 2)) sl.dbg[(sl.dbg_slot == 0 ? 4 : 10) + 4] = te;
    }
  }
}

template <int MODE, int SW, class SL>
__global__ void __launch_bounds__(SW * 4, (SW == 64) ? 2 : 1) k_rl_stream(const __grid_constant__ CUtensorMap tmap,
                                                                         const StreamArgs a, const SL sl) {
  constexpr bool SLAB = !std::is_same<SL, NoSlab>::value;
  SlabCore core;
  int seg_row0, rows_out;
  if constexpr (SLAB) {
    const int seg = slab_segment_of_slot(blockIdx.y, gridDim.y);
    seg_row0 = sl.seg_start[seg];
    rows_out = sl.seg_start[seg + 1] - seg_row0;
    core.row_off = sl.row_off; core.halo = sl.halo; core.own = sl.own;
    core.top = seg == 0; core.bottom = seg == (int)gridDim.y - 1;
    core.up_out = sl.up_out; core.down_out = sl.down_out;
    core.wait_up = sl.wait_up; core.wait_down = sl.wait_down; core.wait_val = sl.wait_val;
    core.cnt_up = sl.cnt_up; core.cnt_down = sl.cnt_down; core.cnt_target = sl.cnt_target;
    core.sig_up = sl.sig_up; core.sig_down = sl.sig_down; core.sig_val = sl.sig_val;
    core.err = sl.err;
  } else {
    seg_row0 = blockIdx.y * a.seg_rows;
    rows_out = min(a.seg_rows, a.Hp - seg_row0);
  }
  if (rows_out <= 0) return;
  rl_stream_body<MODE, SW, SLAB>(&tmap, a, core, blockIdx.x, seg_row0, rows_out);
}

// ---- all bands that still iterate, in ONE launch per filtering -----------------------------------------------
// Richardson-Lucy runs once per FIR band with very different iteration counts (423, 251, 127, 46, 13, 4, 3, 1 for
// the shipped PSF and 8 bands).  Band after band that is 2 * sum(n_iter) = 1736 dependent launches; batching
// iteration i of every band with n_iter > i into one launch (blockIdx.z = band) leaves 2 * max(n_iter) = 846, and
// the launches of the small slabs of a multi-GPU run fill the SMs with several bands at once.  Per-band arguments
// (tensor maps included) live in device memory.
struct BandLaunch {
  CUtensorMap map_u, map_r;           // 128-byte objects, the struct is 128-byte aligned
  StreamArgs sa;                      // geometry; wx / wy / d / out are set per filtering in the kernel
  const float* taps[2];               // [orientation] -> wx' (WU + 8) | wy' (KW + 8)
  float *d, *u, *r;
  int strips, nseg, n_iter, uniform_seg_rows;   // uniform_seg_rows > 0: segments of that many rows (unsharded)
  // slab form
  int row_off, halo, own;
  int seg_start[kMaxSlabSegs + 1];
  int nseg_lone;                       // segmentation of the launches in which this band iterates alone
  int seg_start_lone[kMaxSlabSegs + 1];
  float *up_u, *down_u, *up_r, *down_r;
  const unsigned long long *wait_u_up, *wait_u_down, *wait_r_up, *wait_r_down;
  unsigned long long *cnt_u_up, *cnt_u_down, *cnt_r_up, *cnt_r_down;
  unsigned long long *sig_u_up, *sig_u_down, *sig_r_up, *sig_r_down;
  unsigned long long u_base, r_base, cnt_u_base, cnt_r_base;   // values at the start of the run
  int* err;
  unsigned long long* dbg;            // THZ_SLAB_TRACE: [2 * n_iter][4] time stamps of this band's launches, or null
};

// grid = (strips, bands, segment slots): CTAs are dispatched x first, then y, then z, so that the boundary segments
// (slots 0 and 1) of EVERY band start before any interior segment: their rows reach the neighbours while the
// interior is still being filtered, and the neighbours' next launch finds its halos in place
template <int MODE, bool SLAB>
__global__ void __launch_bounds__(256, 2) k_rl_multi(const BandLaunch* __restrict__ bl, int it, int lone) {
  // one GPU: band-major (blockIdx.z = band), so that the CTAs in flight share one band's images in L2
  const BandLaunch& B = bl[SLAB ? blockIdx.y : blockIdx.z];
  const int slot = SLAB ? blockIdx.z : blockIdx.y;
  const bool alt = SLAB && lone != 0;
  const int nseg = alt ? B.nseg_lone : B.nseg;
  if (it >= B.n_iter || (int)blockIdx.x >= B.strips || slot >= nseg) return;
  StreamArgs a = B.sa;
  const float* tp = B.taps[MODE == 1 ? 0 : 1];
  a.wx = tp;
  a.wy = tp + a.WU + 8;
  a.d = B.d;
  a.out = (MODE == 1) ? B.r : B.u;
  SlabCore core;
  int seg_row0, rows_out;
  if constexpr (SLAB) {
    const int seg = slab_segment_of_slot(slot, nseg);
    const int* tab = alt ? B.seg_start_lone : B.seg_start;
    seg_row0 = tab[seg];
    rows_out = tab[seg + 1] - seg_row0;
    core.row_off = B.row_off; core.halo = B.halo; core.own = B.own;
    core.top = seg == 0; core.bottom = seg == nseg - 1;
    const unsigned long long i = (unsigned long long)it;
    if constexpr (MODE == 1) {        // consumes u version u_base + it, publishes r version r_base + it
      core.up_out = B.up_r; core.down_out = B.down_r;
      core.wait_up = B.wait_u_up; core.wait_down = B.wait_u_down; core.wait_val = B.u_base + i;
      core.cnt_up = B.cnt_r_up; core.cnt_down = B.cnt_r_down;
      core.cnt_target = B.cnt_r_base + (i + 1ull) * (unsigned long long)B.strips;
      core.sig_up = B.sig_r_up; core.sig_down = B.sig_r_down; core.sig_val = B.r_base + i;
    } else {                          // consumes r version r_base + it, publishes u version u_base + it + 1
      core.up_out = B.up_u; core.down_out = B.down_u;
      core.wait_up = B.wait_r_up; core.wait_down = B.wait_r_down; core.wait_val = B.r_base + i;
      core.cnt_up = B.cnt_u_up; core.cnt_down = B.cnt_u_down;
      core.cnt_target = B.cnt_u_base + (i + 1ull) * (unsigned long long)B.strips;
      core.sig_up = B.sig_u_up; core.sig_down = B.sig_u_down; core.sig_val = B.u_base + i + 1ull;
    }
    core.err = B.err;
    core.dbg = B.dbg ? B.dbg + (size_t)(2 * it + (MODE == 1 ? 0 : 1)) * 16 : nullptr;
    core.dbg_slot = slot;
  } else {
    seg_row0 = slot * B.uniform_seg_rows;
    rows_out = min(B.uniform_seg_rows, a.Hp - seg_row0);
  }
  if (rows_out <= 0) return;
  rl_stream_body<MODE, 64, SLAB>((MODE == 1) ? &B.map_u : &B.map_r, a, core, blockIdx.x, seg_row0, rows_out);
}

// Start of a slab run: u_0 = d on the own rows; the boundary rows go to the neighbours' halos and the version is
// published (same protocol as k_rl_stream).  One CTA per pushed row.
__global__ void k_slab_push(const float* __restrict__ img, int pitch, int Wp, SlabArgs sl) {
  const int row = blockIdx.x;          // [0, halo): top rows -> rank above; [halo, 2 halo): bottom rows -> rank below
  const bool up = row < sl.halo;
  float* dst = up ? sl.up_out : sl.down_out;
  if (dst) {
    const int orow = up ? row : sl.own - sl.halo + (row - sl.halo);
    const float* src = img + (size_t)(sl.row_off + orow) * pitch;
    float* d = dst + (size_t)(up ? row : row - sl.halo) * pitch;
    for (int c = threadIdx.x; c < Wp; c += blockDim.x) d[c] = src[c];
  }
  if (up) {
    if (sl.sig_up) slab_signal(sl.cnt_up, sl.cnt_target, sl.sig_up, sl.sig_val);
  } else {
    if (sl.sig_down) slab_signal(sl.cnt_down, sl.cnt_target, sl.sig_down, sl.sig_val);
  }
}

// numpy-"reflect" padding of this rank's rows: own row o of the padded domain is padded row own_lo + o
__global__ void k_reflect_pad_slab(const float* __restrict__ img /* this rank's image rows */, int x0, int rows_loc,
                                   int h_total, int w, int pad_y, int pad_x, int own_lo, int own, int row_off,
                                   float* __restrict__ out, int pitch) {
  const int Wp = w + 2 * pad_x;
  const int64_t total = (int64_t)own * Wp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i / Wp), c = (int)(i % Wp);
    const int r = own_lo + o;
    int sr, sc;
    if (r < pad_y) sr = pad_y - r;
    else if (r >= pad_y + h_total) sr = h_total - 2 - (r - pad_y - h_total);
    else sr = r - pad_y;
    if (c < pad_x) sc = pad_x - c;
    else if (c >= pad_x + w) sc = w - 2 - (c - pad_x - w);
    else sc = c - pad_x;
    const int lr = sr - x0;   // inside this rank's slab by construction (slab rows > pad)
    out[(size_t)(row_off + o) * pitch + c] = (lr >= 0 && lr < rows_loc) ? img[(size_t)lr * w + sc] : 0.f;
  }
}

// crop, clamp, gain for this rank's image rows: image row x0 + i is own row (x0 + pad_y - own_lo) + i
__global__ void k_rl_finish_slab(const float* __restrict__ u, int pitch, int first_row, int pad_x, int rows_loc, int w,
                                 const float* __restrict__ d_img, float* __restrict__ deconv, float* __restrict__ gain) {
  const int64_t total = (int64_t)rows_loc * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / w), c = (int)(i % w);
    const float v = fmaxf(u[(size_t)(first_row + r) * pitch + c + pad_x], 0.0f);
    if (deconv) deconv[i] = v;
    if (gain) gain[i] = sqrtf(v / d_img[i]);
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
// `taps_dst` (optional): device memory of at least kConvTapFloats floats that receives the taps instead of the
// shared WS_RL_TAPS workspace (the slab form keeps one set per band alive across runs); it must be idle.
constexpr size_t kConvTapFloats = 4096;
static int make_conv_plan(thz_ctx* c, cudaStream_t s, int Hp, int Wp, int pitch, const float* psf_x, int kx,
                          const float* psf_y, int ky, const float* dense, int direct, ConvPlan& cp,
                          float* taps_dst = nullptr) {
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
  void* dp = taps_dst;
  if (taps_dst) {
    if (w.size() > kConvTapFloats) return set_err(c, THZ_EINVAL, "PSF taps do not fit the per-band tap store");
  } else {
    int rc = ws_get(c, WS_RL_TAPS, w.size() * sizeof(float), &dp);
    if (rc != THZ_OK) return rc;
  }
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
      cudaError_t e2 = ensure_dynamic_smem(c, (const void*)kernel, cp.ssmem);
      if (e2 != cudaSuccess) return e2;
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
      return cudaLaunchKernelEx(&cfg, kernel, map, sa, NoSlab{});
    };
    e = (cp.sw == 64) ? launch(k_rl_stream<MODE, 64, NoSlab>, 256) : launch(k_rl_stream<MODE, 128, NoSlab>, 512);
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
      e = ensure_dynamic_smem(c, (const void*)k_rl_conv_persistent<MODE>, smem);
      if (e != cudaSuccess) return cuda_fail(c, e, "cudaFuncSetAttribute(rl persistent)");
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
  e = ensure_dynamic_smem(c, key, cp.smem);
  if (e != cudaSuccess) return cuda_fail(c, e, "cudaFuncSetAttribute(rl)");
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


// one filtering of every band that still iterates (k_rl_multi); `active` = bands with n_iter > it
template <int MODE, bool SLAB>
static int launch_multi(thz_ctx* c, cudaStream_t s, const BandLaunch* d_bl, int active, int grid_x, int grid_y,
                        size_t smem, int it, int lone = 0) {
  auto kernel = k_rl_multi<MODE, SLAB>;
  cudaError_t e = ensure_dynamic_smem(c, (const void*)kernel, smem);
  if (e != cudaSuccess) return cuda_fail(c, e, "cudaFuncSetAttribute(k_rl_multi)");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = SLAB ? dim3(grid_x, active, grid_y) : dim3(grid_x, grid_y, active);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kernel, d_bl, it, lone);
  c->launches++;
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "k_rl_multi launch");
  return THZ_OK;
}

static void fill_band_launch_common(BandLaunch& L, const ConvPlan& cp, const CUtensorMap& map_u, const CUtensorMap& map_r,
                                    float* d, float* u, float* r, int n_iter) {
  memset(&L, 0, sizeof L);
  L.map_u = map_u;
  L.map_r = map_r;
  L.sa = cp.sa;
  L.taps[0] = cp.d_ws;
  L.taps[1] = cp.d_ws + cp.swstride;
  L.d = d; L.u = u; L.r = r;
  L.strips = (int)cp.sgrid.x;
  L.n_iter = n_iter;
}

// `richardson_lucy` + clamp + gain for ALL bands of a deconvolution on one GPU, iteration i of every band that
// still iterates batched into one launch per filtering.  Returns THZ_SKIP_NO_PSF (never an error of its own)
// when a band's PSF does not run on the 64-column streaming kernel: the caller then iterates band after band.
int richardson_lucy_bands(thz_ctx* c, cudaStream_t s, const float* d_energy, int64_t P, int rows, int cols,
                          const thz_band_plan* bands, int B, float* d_gain, const volatile uint8_t* abort_flag,
                          thz_progress_fn progress, void* puser, long* iterations_run) {
  if (B < 1 || B > THZ_MAX_BANDS) return set_err(c, THZ_EINVAL, "bad band count");
  struct Geo { int pad_y, pad_x, Hp, Wp, pitch; size_t img_bytes; };
  std::vector<Geo> g((size_t)B);
  size_t total = 0;
  for (int b = 0; b < B; ++b) {
    const int kx = bands[b].kx, ky = bands[b].ky;
    if (kx < 1 || ky < 1 || !(kx & 1) || !(ky & 1) || kx > THZ_MAX_PSF || ky > THZ_MAX_PSF) return THZ_SKIP_NO_PSF;
    g[b].pad_y = kx / 2; g[b].pad_x = ky / 2;
    if (g[b].pad_y >= rows - 1 || g[b].pad_x >= cols - 1) return set_err(c, THZ_EINVAL, "PSF larger than the image");
    g[b].Hp = rows + 2 * g[b].pad_y; g[b].Wp = cols + 2 * g[b].pad_x; g[b].pitch = round_up(g[b].Wp, 4);
    g[b].img_bytes = ((size_t)g[b].Hp * g[b].pitch * sizeof(float) + 1023) & ~(size_t)1023;
    total += 3 * g[b].img_bytes + kConvTapFloats * sizeof(float);
  }
  const size_t launch_bytes = ((size_t)B * sizeof(BandLaunch) + 1023) & ~(size_t)1023;
  void* ws = nullptr;
  int rc = ws_get(c, WS_RL_MULTI, total + launch_bytes, &ws);
  if (rc != THZ_OK) return rc;
  unsigned char* base = (unsigned char*)ws;
  BandLaunch* d_bl = reinterpret_cast<BandLaunch*>(base);
  unsigned char* cur = base + launch_bytes;
  std::vector<BandLaunch> h((size_t)B);
  std::vector<int> order((size_t)B);
  for (int b = 0; b < B; ++b) order[b] = b;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return bands[x].n_iter > bands[y].n_iter; });
  struct Buf { float *d, *u, *r; };
  std::vector<Buf> bufs((size_t)B);
  size_t smem = 0;
  int grid_x = 1, grid_y = 1, max_iter = 0;
  for (int b = 0; b < B; ++b) {
    float* taps = reinterpret_cast<float*>(cur);
    cur += kConvTapFloats * sizeof(float);
    bufs[b].d = reinterpret_cast<float*>(cur); cur += g[b].img_bytes;
    bufs[b].u = reinterpret_cast<float*>(cur); cur += g[b].img_bytes;
    bufs[b].r = reinterpret_cast<float*>(cur); cur += g[b].img_bytes;
  }
  THZ_CUDA(c, cudaStreamSynchronize(s));   // nothing of a previous run still reads the taps / launch table
  for (int z = 0; z < B; ++z) {
    const int b = order[z];
    ConvPlan cp;
    float* taps = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bufs[b].d) - kConvTapFloats * sizeof(float));
    rc = make_conv_plan(c, s, g[b].Hp, g[b].Wp, g[b].pitch, bands[b].psf_x, bands[b].kx, bands[b].psf_y, bands[b].ky,
                        nullptr, bands[b].direct ? 1 : 0, cp, taps);
    if (rc != THZ_OK) return rc;
    if (!cp.streaming || cp.sw != 64) return THZ_SKIP_NO_PSF;
    CUtensorMap map_u, map_r;
    rc = make_tmap(c, &map_u, bufs[b].u, g[b].Hp, g[b].Wp, g[b].pitch, kCR, cp.sa.bc);
    if (rc == THZ_OK) rc = make_tmap(c, &map_r, bufs[b].r, g[b].Hp, g[b].Wp, g[b].pitch, kCR, cp.sa.bc);
    if (rc != THZ_OK) return rc;
    fill_band_launch_common(h[z], cp, map_u, map_r, bufs[b].d, bufs[b].u, bufs[b].r, bands[b].n_iter);
    h[z].nseg = (int)cp.sgrid.y;
    h[z].uniform_seg_rows = cp.sa.seg_rows;
    smem = std::max(smem, cp.ssmem);
    grid_x = std::max(grid_x, h[z].strips);
    grid_y = std::max(grid_y, h[z].nseg);
    max_iter = std::max(max_iter, bands[b].n_iter);
    // d = reflect-padded band image, u_0 = d, r = 0 (its pitch padding and the rows TMA never writes stay 0)
    THZ_CUDA(c, cudaMemsetAsync(bufs[b].d, 0, g[b].img_bytes, s));
    k_reflect_pad<<<grid_for(c, (int64_t)g[b].Hp * g[b].Wp), 256, 0, s>>>(d_energy + (size_t)b * P, rows, cols, g[b].pad_y,
                                                                       g[b].pad_x, bufs[b].d, g[b].pitch);
    c->launches++;
    THZ_CUDA(c, cudaMemcpyAsync(bufs[b].u, bufs[b].d, g[b].img_bytes, cudaMemcpyDeviceToDevice, s));
    THZ_CUDA(c, cudaMemsetAsync(bufs[b].r, 0, g[b].img_bytes, s));
  }
  THZ_CUDA(c, cudaMemcpyAsync(d_bl, h.data(), (size_t)B * sizeof(BandLaunch), cudaMemcpyHostToDevice, s));
  long total_iter = 0, done_iter = 0;
  for (int b = 0; b < B; ++b) total_iter += std::max(bands[b].n_iter, 1);
  bool aborted = false;
  for (int it = 0; rc == THZ_OK && it < max_iter; ++it) {
    int active = 0;
    while (active < B && h[active].n_iter > it) ++active;
    rc = launch_multi<1, false>(c, s, d_bl, active, grid_x, grid_y, smem, it);          // r = d / (u (*) psf + eps)
    if (rc == THZ_OK) rc = launch_multi<2, false>(c, s, d_bl, active, grid_x, grid_y, smem, it);   // u *= r (*) mirror
    done_iter += active;
    if ((it & 15) == 15 || it == max_iter - 1) {
      if (abort_flag && *abort_flag) { aborted = true; break; }
      if (progress || abort_flag) {
        cudaError_t e = cudaStreamSynchronize(s);   // keep the queue short so that abort is responsive
        if (e != cudaSuccess) { rc = cuda_fail(c, e, "richardson_lucy_bands"); break; }
        if (progress) progress(0.1f + 0.8f * (float)done_iter / (float)total_iter, puser);
      }
    }
  }
  if (rc == THZ_OK && !aborted)
    for (int b = 0; b < B; ++b) {
      k_rl_finish<<<grid_for(c, (int64_t)rows * cols), 256, 0, s>>>(bufs[b].u, g[b].pitch, g[b].pad_y, g[b].pad_x, rows,
                                                                    cols, d_energy + (size_t)b * P, nullptr,
                                                                    d_gain + (size_t)b * P);
      c->launches++;
    }
  if (iterations_run) *iterations_run = done_iter;
  cudaError_t e = cudaStreamSynchronize(s);
  if (rc == THZ_OK && e != cudaSuccess) return cuda_fail(c, e, "richardson_lucy_bands");
  if (rc == THZ_OK && aborted) return THZ_ABORTED;
  return rc;
}

// ------------------------------------------------------------------------------------
// Row-slab Richardson-Lucy over several GPUs: every rank iterates its own rows of the padded
// domain; the kx/2 boundary rows travel as peer stores inside the filtering kernels (see SlabArgs).
//
// Arena: one cudaMalloc per rank, identical layout on every rank, mapped into the two neighbours (same process:
// plain peer pointers; one process per GPU: cudaIpc handles exchanged by the host):
//   [0, 1 KiB)            error word
//   per band b:           flags (8 x 128 B) | taps (16 KiB) | d | u | r, images of (halo_b + own_max + halo_b) rows
// ------------------------------------------------------------------------------------
constexpr size_t kSlabFlagBytes = 8 * 128;
constexpr size_t kSlabTapBytes = kConvTapFloats * sizeof(float);

struct SlabBand {
  int kx = 0, ky = 0, direct = 0, n_iter = 0;
  int halo = 0, pad_x = 0;
  int Hp = 0, Wp = 0, pitch = 0;
  int own_lo = 0, own = 0;          // own rows of the padded domain [own_lo, own_lo + own)
  int own_up = 0;                   // own rows of the rank above (its bottom halo starts at row halo + own_up)
  int rows_loc = 0;                 // halo + own + halo
  size_t off_flags = 0, off_taps = 0, off_d = 0, off_u = 0, off_r = 0, img_bytes = 0;
  ConvPlan cp;
  CUtensorMap map_u, map_r;
  int nseg = 0, strips = 0;
  int seg_start[kMaxSlabSegs + 1] = {};      // few long segments: launches that several bands share
  int nseg_lone = 0;
  int seg_start_lone[kMaxSlabSegs + 1] = {}; // many short segments: launches in which this band iterates alone
  unsigned long long u_next = 1, r_next = 1;   // next version numbers of this rank's u / r boundary rows
  unsigned long long cnt_u = 0, cnt_r = 0;     // cumulative boundary-CTA counts (both directions advance alike)
};

}  // namespace thz

struct thz_slab {
  thz_ctx* ctx = nullptr;
  int rank = 0, world = 1;
  std::vector<int> bounds;          // image row bounds of every rank, [world + 1]
  int cols = 0;
  unsigned char* arena = nullptr;
  size_t arena_bytes = 0;
  unsigned char* up = nullptr;      // neighbours' arenas as seen from this rank
  unsigned char* down = nullptr;
  bool up_ipc = false, down_ipc = false;
  bool connected = false;
  std::vector<thz::SlabBand> bands;
  std::vector<unsigned char> key;   // geometry + PSF bytes of the current plan
  cudaStream_t stream = nullptr;    // stream the run is launched on (the context's by default)
  void* d_launch = nullptr;         // BandLaunch table of the batched kernels
  bool trace = false;               // THZ_SLAB_TRACE=<file prefix>: per-launch time stamps of the first band
  std::string trace_path;
  void* d_trace = nullptr;
  size_t d_trace_bytes = 0;
  int trace_launches = 0;
};

namespace thz {

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static unsigned long long* slab_flag(unsigned char* arena, const SlabBand& b, int idx) {
  return reinterpret_cast<unsigned long long*>(arena + b.off_flags + (size_t)idx * 128);
}
enum { F_U_FROM_UP = 0, F_U_FROM_DOWN, F_R_FROM_UP, F_R_FROM_DOWN, F_CNT_U_UP, F_CNT_U_DOWN, F_CNT_R_UP, F_CNT_R_DOWN };

// geometry of band `pb` on this rank; layout offsets are the same on every rank
static int slab_band_geometry(thz_slab* sl, const thz_band_plan& pb, size_t& cursor, SlabBand& b) {
  thz_ctx* c = sl->ctx;
  const int rows_total = sl->bounds[sl->world], w = sl->cols;
  b.kx = pb.kx; b.ky = pb.ky; b.direct = pb.direct; b.n_iter = pb.n_iter;
  b.halo = pb.kx / 2;               // psf.nrows() / 2 pads axis 0 (deconvolution.rs:629-631)
  b.pad_x = pb.ky / 2;
  if (b.halo >= rows_total - 1 || b.pad_x >= w - 1) return set_err(c, THZ_EINVAL, "PSF larger than the image");
  b.Hp = rows_total + 2 * b.halo;
  b.Wp = w + 2 * b.pad_x;
  b.pitch = round_up(b.Wp, 4);
  int own_max = 0;
  for (int r = 0; r < sl->world; ++r) {
    int lo = sl->bounds[r] + b.halo, hi = sl->bounds[r + 1] + b.halo;
    if (r == 0) lo = 0;
    if (r == sl->world - 1) hi = b.Hp;
    const int own = hi - lo;
    // the reflected pad rows come from this rank's own slab, and a boundary segment must not reach the other one
    if (sl->bounds[r + 1] - sl->bounds[r] <= b.halo || own < 3 * std::max(b.halo, 1))
      return set_err(c, THZ_EINVAL, "row slab thinner than three PSF half-heights: use fewer ranks for this image");
    own_max = std::max(own_max, own);
    if (r == sl->rank) { b.own_lo = lo; b.own = own; }
    if (r == sl->rank - 1) b.own_up = own;
  }
  b.rows_loc = b.own + 2 * b.halo;
  b.img_bytes = align_up((size_t)(own_max + 2 * b.halo) * b.pitch * sizeof(float), 1024);
  b.off_flags = cursor;
  b.off_taps = b.off_flags + kSlabFlagBytes;
  b.off_d = align_up(b.off_taps + kSlabTapBytes, 1024);
  b.off_u = b.off_d + b.img_bytes;
  b.off_r = b.off_u + b.img_bytes;
  cursor = b.off_r + b.img_bytes;
  return THZ_OK;
}

static int slab_band_finish_plan(thz_slab* sl, const thz_band_plan& pb, SlabBand& b) {
  thz_ctx* c = sl->ctx;
  int rc = make_conv_plan(c, sl->stream, b.rows_loc, b.Wp, b.pitch, pb.psf_x, pb.kx, pb.psf_y, pb.ky, nullptr,
                          pb.direct ? 1 : 0, b.cp, reinterpret_cast<float*>(sl->arena + b.off_taps));
  if (rc != THZ_OK) return rc;
  if (!b.cp.streaming) return set_err(c, THZ_EINVAL, "PSF too large for the streaming kernel: slab form unavailable");
  rc = make_tmap(c, &b.map_u, reinterpret_cast<float*>(sl->arena + b.off_u), b.rows_loc, b.Wp, b.pitch, kCR, b.cp.sa.bc);
  if (rc == THZ_OK)
    rc = make_tmap(c, &b.map_r, reinterpret_cast<float*>(sl->arena + b.off_r), b.rows_loc, b.Wp, b.pitch, kCR, b.cp.sa.bc);
  if (rc != THZ_OK) return rc;
  // segments: top boundary | interior ... | bottom boundary (halo = 0: a 1 x ky PSF needs no exchange at all)
  const StreamArgs& sa = b.cp.sa;
  b.strips = (int)b.cp.sgrid.x;
  const int h = std::max(b.halo, 1);
  const int per_sm = (b.cp.sw == 64 ? 2 : 1);
  // Segments.  A CTA works in 64-row chunks and re-filters WU warm-up rows, so a segment of 64 c - WU rows uses its
  // c chunks fully.  The boundary segments must hold the halo rows (they read the neighbours' rows and produce the
  // rows the neighbours need) but may be longer: a 23-row boundary costs two chunks either way, an 80-row one does
  // 57 interior rows in the same time.  Plans: every segment c chunks long, the boundaries included; c is chosen
  // per table from measured costs (profiles/r02_slab_trace_*.txt: ~6 us per chunk when a CTA has its SM to itself,
  // 3.25 us per chunk and CTA when two interleave; a launch lasts as long as its busiest SM):
  //   lone  : min over c of  m x c x cost(m),  m = CTAs per SM -- launches in which this band iterates alone
  //   shared: CTAs <= SMs with the fewest chunk units -- launches several bands share fill each other's SMs
  auto chunks_of = [&](int rows) { return (rows + sa.WU + kCR - 1) / kCR; };
  struct Cand { int b_rows, seg_rows, ni, ctas, c; double t; };
  const int sms = c->sm_count;
  auto make = [&](int nch) -> Cand {
    const int c = nch;
    Cand k{};
    k.c = c;
    k.b_rows = c * kCR - sa.WU;
    if (k.b_rows < h || 2 * k.b_rows > b.own) {   // too short for the halo / too long for the slab: tight boundaries
      k.b_rows = h;
      k.c = std::max(c, chunks_of(h));
    }
    const int interior = b.own - 2 * k.b_rows;
    k.seg_rows = std::max(1, c * kCR - sa.WU);
    k.ni = interior > 0 ? (interior + k.seg_rows - 1) / k.seg_rows : 0;
    if (k.ni > 0) k.seg_rows = (interior + k.ni - 1) / k.ni;   // equal parts
    k.ctas = b.strips * (k.ni + 2);
    const int m = (k.ctas + sms - 1) / sms;
    k.t = (double)m * k.c * (m == 1 ? 6.0 : 3.25);
    return k;
  };
  Cand lone{}, shared{};
  lone.t = 1e30;
  double shared_units = 1e30;
  bool have_shared = false;
  for (int cc = 1; cc <= 40; ++cc) {
    const Cand k = make(cc);
    if (k.ni + 2 > kMaxSlabSegs) continue;
    if (k.ctas <= sms * per_sm && k.t < lone.t - 1e-9) lone = k;
    if (k.ctas <= sms && (double)k.ctas * k.c < shared_units - 1e-9) {
      shared_units = (double)k.ctas * k.c;
      shared = k;
      have_shared = true;
    }
    if (2 * (cc * kCR - sa.WU) > b.own && cc > chunks_of(h)) break;
  }
  if (lone.t >= 1e30) lone = make(std::max(1, chunks_of(b.own / 2)));   // cannot happen for sane sizes
  if (!have_shared) shared = lone;
  auto build = [&](const Cand& k, int* seg, int& nseg) {
    nseg = 0;
    seg[nseg++] = 0;
    seg[nseg++] = k.b_rows;
    for (int i = 1; i < k.ni; ++i) seg[nseg++] = k.b_rows + i * k.seg_rows;
    if (k.ni > 0) seg[nseg++] = b.own - k.b_rows;
    seg[nseg] = b.own;
  };
  build(shared, b.seg_start, b.nseg);
  build(lone, b.seg_start_lone, b.nseg_lone);
  return THZ_OK;
}

static void slab_fill_args(const thz_slab* sl, const SlabBand& b, SlabArgs& a) {
  a = SlabArgs{};
  a.row_off = b.halo;
  a.halo = b.halo;
  a.own = b.own;
  for (int i = 0; i <= b.nseg_lone; ++i) a.seg_start[i] = b.seg_start_lone[i];
  a.err = reinterpret_cast<int*>(sl->arena);
}

// image `which` (off_u / off_r) of band b: where this rank's boundary rows land in the neighbours
static float* slab_up_dst(const thz_slab* sl, const SlabBand& b, size_t off_img) {
  if (!sl->up || b.halo == 0) return nullptr;
  return reinterpret_cast<float*>(sl->up + off_img) + (size_t)(b.halo + b.own_up) * b.pitch;   // its bottom halo
}
static float* slab_down_dst(const thz_slab* sl, const SlabBand& b, size_t off_img) {
  if (!sl->down || b.halo == 0) return nullptr;
  return reinterpret_cast<float*>(sl->down + off_img);                                         // its top halo
}

template <int MODE>
static int slab_launch_conv(thz_slab* sl, SlabBand& b, cudaStream_t s, unsigned long long wait_val,
                            unsigned long long sig_val) {
  thz_ctx* c = sl->ctx;
  StreamArgs sa = b.cp.sa;
  const int orient = (MODE == 1) ? 0 : 1;
  sa.wx = b.cp.d_ws + (size_t)orient * b.cp.swstride;
  sa.wy = sa.wx + sa.WU + 8;
  float* d = reinterpret_cast<float*>(sl->arena + b.off_d);
  float* u = reinterpret_cast<float*>(sl->arena + b.off_u);
  float* r = reinterpret_cast<float*>(sl->arena + b.off_r);
  sa.d = d;
  sa.out = (MODE == 1) ? r : u;
  SlabArgs a;
  slab_fill_args(sl, b, a);
  const size_t off_out = (MODE == 1) ? b.off_r : b.off_u;
  a.up_out = slab_up_dst(sl, b, off_out);
  a.down_out = slab_down_dst(sl, b, off_out);
  // MODE 1 reads u and publishes r; MODE 2 reads r and publishes u
  const bool has_up = sl->up && b.halo > 0, has_down = sl->down && b.halo > 0;
  a.wait_up = has_up ? slab_flag(sl->arena, b, MODE == 1 ? F_U_FROM_UP : F_R_FROM_UP) : nullptr;
  a.wait_down = has_down ? slab_flag(sl->arena, b, MODE == 1 ? F_U_FROM_DOWN : F_R_FROM_DOWN) : nullptr;
  a.wait_val = wait_val;
  a.cnt_up = slab_flag(sl->arena, b, MODE == 1 ? F_CNT_R_UP : F_CNT_U_UP);
  a.cnt_down = slab_flag(sl->arena, b, MODE == 1 ? F_CNT_R_DOWN : F_CNT_U_DOWN);
  unsigned long long& cnt = (MODE == 1) ? b.cnt_r : b.cnt_u;
  cnt += (unsigned long long)b.strips;
  a.cnt_target = cnt;
  // the rank above reads its "from down" flag, the rank below its "from up" flag
  a.sig_up = has_up ? slab_flag(sl->up, b, MODE == 1 ? F_R_FROM_DOWN : F_U_FROM_DOWN) : nullptr;
  a.sig_down = has_down ? slab_flag(sl->down, b, MODE == 1 ? F_R_FROM_UP : F_U_FROM_UP) : nullptr;
  a.sig_val = sig_val;
  const CUtensorMap& map = (MODE == 1) ? b.map_u : b.map_r;
  auto launch = [&](auto kernel, int threads) -> cudaError_t {
    cudaError_t e2 = ensure_dynamic_smem(c, (const void*)kernel, b.cp.ssmem);
    if (e2 != cudaSuccess) return e2;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(b.strips, b.nseg_lone);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = b.cp.ssmem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, map, sa, a);
  };
  cudaError_t e = (b.cp.sw == 64) ? launch(k_rl_stream<MODE, 64, SlabArgs>, 256)
                                  : launch(k_rl_stream<MODE, 128, SlabArgs>, 512);
  c->launches++;
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "k_rl_stream (slab) launch");
  return THZ_OK;
}

// run start of band b on this rank: d = reflect pad of the rank's energy rows, u_0 = d, boundary rows pushed
static int slab_band_begin(thz_slab* sl, SlabBand& b, cudaStream_t s, const float* d_energy_b) {
  thz_ctx* c = sl->ctx;
  float* d = reinterpret_cast<float*>(sl->arena + b.off_d);
  float* u = reinterpret_cast<float*>(sl->arena + b.off_u);
  const int x0 = sl->bounds[sl->rank], rows_loc = sl->bounds[sl->rank + 1] - x0;
  k_reflect_pad_slab<<<grid_for(c, (int64_t)b.own * b.Wp), 256, 0, s>>>(d_energy_b, x0, rows_loc, sl->bounds[sl->world],
                                                                       sl->cols, b.halo, b.pad_x, b.own_lo, b.own,
                                                                       b.halo, d, b.pitch);
  c->launches++;
  // own rows only: the halo rows belong to the neighbours, who may already have written their next version
  THZ_CUDA(c, cudaMemcpyAsync(u + (size_t)b.halo * b.pitch, d + (size_t)b.halo * b.pitch,
                              (size_t)b.own * b.pitch * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (b.halo > 0 && (sl->up || sl->down)) {
    SlabArgs a;
    slab_fill_args(sl, b, a);
    a.up_out = slab_up_dst(sl, b, b.off_u);
    a.down_out = slab_down_dst(sl, b, b.off_u);
    a.cnt_up = slab_flag(sl->arena, b, F_CNT_U_UP);
    a.cnt_down = slab_flag(sl->arena, b, F_CNT_U_DOWN);
    b.cnt_u += (unsigned long long)b.halo;
    a.cnt_target = b.cnt_u;
    a.sig_up = sl->up ? slab_flag(sl->up, b, F_U_FROM_DOWN) : nullptr;
    a.sig_down = sl->down ? slab_flag(sl->down, b, F_U_FROM_UP) : nullptr;
    a.sig_val = b.u_next;
    k_slab_push<<<2 * b.halo, 256, 0, s>>>(u, b.pitch, b.Wp, a);
    c->launches++;
  } else {
    b.cnt_u += (unsigned long long)b.halo;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "slab run start");
  return THZ_OK;
}

static int slab_band_end(thz_slab* sl, SlabBand& b, cudaStream_t s, const float* d_energy_b, float* d_deconv_b,
                         float* d_gain_b) {
  thz_ctx* c = sl->ctx;
  const int x0 = sl->bounds[sl->rank], rows_loc = sl->bounds[sl->rank + 1] - x0;
  const float* u = reinterpret_cast<const float*>(sl->arena + b.off_u);
  // image row x0 is padded row x0 + halo = own row x0 + halo - own_lo, buffer row halo + that
  const int first_row = b.halo + (x0 + b.halo - b.own_lo);
  k_rl_finish_slab<<<grid_for(c, (int64_t)rows_loc * sl->cols), 256, 0, s>>>(u, b.pitch, first_row, b.pad_x, rows_loc,
                                                                            sl->cols, d_energy_b, d_deconv_b, d_gain_b);
  c->launches++;
  b.u_next += (unsigned long long)b.n_iter + 1ull;
  b.r_next += (unsigned long long)b.n_iter;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "slab run end");
  return THZ_OK;
}

// launch table of a slab run: one BandLaunch per band, bands sorted by falling iteration count; the version /
// counter bases are those at the start of the run, after the start-of-run push
static void slab_fill_launch(thz_slab* sl, std::vector<BandLaunch>& h, std::vector<int>& order) {
  const int B = (int)sl->bands.size();
  order.resize((size_t)B);
  for (int b = 0; b < B; ++b) order[b] = b;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return sl->bands[x].n_iter > sl->bands[y].n_iter; });
  h.resize((size_t)B);
  for (int z = 0; z < B; ++z) {
    SlabBand& b = sl->bands[order[z]];
    BandLaunch& L = h[z];
    fill_band_launch_common(L, b.cp, b.map_u, b.map_r, reinterpret_cast<float*>(sl->arena + b.off_d),
                            reinterpret_cast<float*>(sl->arena + b.off_u), reinterpret_cast<float*>(sl->arena + b.off_r),
                            b.n_iter);
    L.strips = b.strips;
    L.nseg = b.nseg;
    L.row_off = b.halo; L.halo = b.halo; L.own = b.own;
    for (int i = 0; i <= b.nseg; ++i) L.seg_start[i] = b.seg_start[i];
    L.nseg_lone = b.nseg_lone;
    for (int i = 0; i <= b.nseg_lone; ++i) L.seg_start_lone[i] = b.seg_start_lone[i];
    const bool has_up = sl->up && b.halo > 0, has_down = sl->down && b.halo > 0;
    L.up_u = slab_up_dst(sl, b, b.off_u); L.down_u = slab_down_dst(sl, b, b.off_u);
    L.up_r = slab_up_dst(sl, b, b.off_r); L.down_r = slab_down_dst(sl, b, b.off_r);
    L.wait_u_up = has_up ? slab_flag(sl->arena, b, F_U_FROM_UP) : nullptr;
    L.wait_u_down = has_down ? slab_flag(sl->arena, b, F_U_FROM_DOWN) : nullptr;
    L.wait_r_up = has_up ? slab_flag(sl->arena, b, F_R_FROM_UP) : nullptr;
    L.wait_r_down = has_down ? slab_flag(sl->arena, b, F_R_FROM_DOWN) : nullptr;
    L.cnt_u_up = slab_flag(sl->arena, b, F_CNT_U_UP); L.cnt_u_down = slab_flag(sl->arena, b, F_CNT_U_DOWN);
    L.cnt_r_up = slab_flag(sl->arena, b, F_CNT_R_UP); L.cnt_r_down = slab_flag(sl->arena, b, F_CNT_R_DOWN);
    L.sig_u_up = has_up ? slab_flag(sl->up, b, F_U_FROM_DOWN) : nullptr;
    L.sig_u_down = has_down ? slab_flag(sl->down, b, F_U_FROM_UP) : nullptr;
    L.sig_r_up = has_up ? slab_flag(sl->up, b, F_R_FROM_DOWN) : nullptr;
    L.sig_r_down = has_down ? slab_flag(sl->down, b, F_R_FROM_UP) : nullptr;
    L.u_base = b.u_next; L.r_base = b.r_next;
    L.cnt_u_base = b.cnt_u; L.cnt_r_base = b.cnt_r;
    L.err = reinterpret_cast<int*>(sl->arena);
    L.dbg = nullptr;
    if (z == 0 && sl->trace) {      // time stamps of the longest-running band's launches
      const size_t bytes = (size_t)2 * b.n_iter * 16 * sizeof(unsigned long long);
      if (sl->d_trace_bytes < bytes) {
        if (sl->d_trace) cudaFree(sl->d_trace);
        sl->d_trace = nullptr;
        if (cudaMalloc(&sl->d_trace, bytes) == cudaSuccess) sl->d_trace_bytes = bytes;
      }
      if (sl->d_trace) {
        std::vector<unsigned long long> init((size_t)2 * b.n_iter * 16, 0ull);
        for (size_t i = 0; i < init.size(); i += 16) init[i] = ~0ull;
        cudaMemcpyAsync(sl->d_trace, init.data(), bytes, cudaMemcpyHostToDevice, sl->stream);
        L.dbg = (unsigned long long*)sl->d_trace;
        sl->trace_launches = 2 * b.n_iter;
      }
    }
  }
}

// One run over all bands for a set of ranks that live in this process.  `nranks` > 1 = the ranks share one GPU
// (the single-GPU emulation used by the tests): every kernel goes to ONE stream in dependency order, so that no
// kernel ever waits for one that has not been launched ahead of it.
int slab_run(thz_slab* const* ranks, int nranks, const float* const* d_energy, const int64_t* bstride,
             float* const* d_gain, float* const* d_deconv) {
  if (nranks < 1) return THZ_EINVAL;
  const size_t B = ranks[0]->bands.size();
  for (int r = 0; r < nranks; ++r) {
    thz_slab* sl = ranks[r];
    if (!sl->connected) return set_err(sl->ctx, THZ_ESTATE, "thz_slab_connect_* has not been called for the current plan");
    if (sl->bands.size() != B) return set_err(sl->ctx, THZ_ESTATE, "ranks hold different plans");
  }
  int rc = THZ_OK;
  // start of the run: every band's d, u_0 and the first push of boundary rows
  for (size_t b = 0; rc == THZ_OK && b < B; ++b)
    for (int r = 0; rc == THZ_OK && r < nranks; ++r) {
      thz_slab* sl = ranks[r];
      cudaSetDevice(sl->ctx->device);
      rc = slab_band_begin(sl, sl->bands[b], sl->stream, d_energy[r] + b * (size_t)bstride[r]);
    }
  bool batched = ranks[0]->ctx->rl_batch;
  for (int r = 0; r < nranks; ++r)
    for (const SlabBand& b : ranks[r]->bands) batched = batched && b.cp.sw == 64;
  if (rc == THZ_OK && batched) {
    int max_iter = 0;
    std::vector<std::vector<BandLaunch>> h((size_t)nranks);
    std::vector<int> order;
    struct G { int gx = 1, gy = 1, gy_lone = 1; size_t smem = 0; };
    std::vector<G> g((size_t)nranks);
    for (int r = 0; rc == THZ_OK && r < nranks; ++r) {
      thz_slab* sl = ranks[r];
      cudaSetDevice(sl->ctx->device);
      slab_fill_launch(sl, h[r], order);
      for (const BandLaunch& L : h[r]) {
        g[r].gx = std::max(g[r].gx, L.strips);
        g[r].gy = std::max(g[r].gy, L.nseg);
        max_iter = std::max(max_iter, L.n_iter);
      }
      g[r].gy_lone = h[r][0].nseg_lone;      // the band with the most iterations is the one that ends up alone
      for (const SlabBand& b : sl->bands) g[r].smem = std::max(g[r].smem, b.cp.ssmem);
      if (!sl->d_launch) {
        void* p = nullptr;
        THZ_CUDA(sl->ctx, cudaMalloc(&p, THZ_MAX_BANDS * sizeof(BandLaunch)));
        sl->d_launch = p;
      }
      THZ_CUDA(sl->ctx, cudaMemcpyAsync(sl->d_launch, h[r].data(), B * sizeof(BandLaunch), cudaMemcpyHostToDevice,
                                        sl->stream));
    }
    for (int it = 0; rc == THZ_OK && it < max_iter; ++it) {
      int active = 0;
      while (active < (int)B && h[0][active].n_iter > it) ++active;
      const int lone = active == 1 ? 1 : 0;
      for (int r = 0; rc == THZ_OK && r < nranks; ++r)
        rc = launch_multi<1, true>(ranks[r]->ctx, ranks[r]->stream, (const BandLaunch*)ranks[r]->d_launch, active, g[r].gx,
                                   lone ? g[r].gy_lone : g[r].gy, g[r].smem, it, lone);
      for (int r = 0; rc == THZ_OK && r < nranks; ++r)
        rc = launch_multi<2, true>(ranks[r]->ctx, ranks[r]->stream, (const BandLaunch*)ranks[r]->d_launch, active, g[r].gx,
                                   lone ? g[r].gy_lone : g[r].gy, g[r].smem, it, lone);
    }
    // the counters the kernels advanced
    for (int r = 0; r < nranks; ++r)
      for (SlabBand& b : ranks[r]->bands) {
        b.cnt_u += (unsigned long long)b.n_iter * (unsigned long long)b.strips;
        b.cnt_r += (unsigned long long)b.n_iter * (unsigned long long)b.strips;
      }
  } else {
    for (size_t b = 0; rc == THZ_OK && b < B; ++b) {
      const int n_iter = ranks[0]->bands[b].n_iter;
      for (int it = 0; rc == THZ_OK && it < n_iter; ++it) {
        for (int r = 0; rc == THZ_OK && r < nranks; ++r) {
          thz_slab* sl = ranks[r];
          SlabBand& bb = sl->bands[b];
          rc = slab_launch_conv<1>(sl, bb, sl->stream, bb.u_next + (unsigned long long)it, bb.r_next + (unsigned long long)it);
        }
        for (int r = 0; rc == THZ_OK && r < nranks; ++r) {
          thz_slab* sl = ranks[r];
          SlabBand& bb = sl->bands[b];
          rc = slab_launch_conv<2>(sl, bb, sl->stream, bb.r_next + (unsigned long long)it,
                                   bb.u_next + (unsigned long long)it + 1ull);
        }
      }
    }
  }
  for (size_t b = 0; rc == THZ_OK && b < B; ++b)
    for (int r = 0; rc == THZ_OK && r < nranks; ++r) {
      thz_slab* sl = ranks[r];
      cudaSetDevice(sl->ctx->device);
      rc = slab_band_end(sl, sl->bands[b], sl->stream, d_energy[r] + b * (size_t)bstride[r],
                         d_deconv && d_deconv[r] ? d_deconv[r] + b * (size_t)bstride[r] : nullptr,
                         d_gain[r] + b * (size_t)bstride[r]);
    }
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


// ------------------------------------------------------------------------- row-slab Richardson-Lucy
int thz_slab_create(thz_ctx* c, int rank, int world, thz_slab** out) {
  CHECK_CTX(c);
  if (!out || world < 1 || rank < 0 || rank >= world) return set_err(c, THZ_EINVAL, "bad rank / world");
  thz_slab* sl = new thz_slab();
  sl->ctx = c;
  sl->rank = rank;
  sl->world = world;
  sl->stream = c->stream;
  if (const char* t = getenv("THZ_SLAB_TRACE")) {
    sl->trace = true;
    sl->trace_path = std::string(t) + ".rank" + std::to_string(rank) + ".csv";
  }
  *out = sl;
  return THZ_OK;
}

static void slab_disconnect(thz_slab* sl) {
  if (sl->up && sl->up_ipc) cudaIpcCloseMemHandle(sl->up);
  if (sl->down && sl->down_ipc) cudaIpcCloseMemHandle(sl->down);
  sl->up = sl->down = nullptr;
  sl->up_ipc = sl->down_ipc = false;
  sl->connected = false;
}

void thz_slab_destroy(thz_slab* sl) {
  if (!sl) return;
  cudaSetDevice(sl->ctx->device);
  cudaStreamSynchronize(sl->stream);
  slab_disconnect(sl);
  if (sl->arena) cudaFree(sl->arena);
  if (sl->d_launch) cudaFree(sl->d_launch);
  if (sl->d_trace) cudaFree(sl->d_trace);
  delete sl;
}

int thz_slab_plan(thz_slab* sl, const int* row_bounds, int cols, const thz_band_plan* bands, int n_bands, int* changed) {
  if (!sl) return THZ_EINVAL;
  thz_ctx* c = sl->ctx;
  CHECK_CTX(c);
  if (!row_bounds || !bands || n_bands < 1 || n_bands > THZ_MAX_BANDS || cols < 2)
    return set_err(c, THZ_EINVAL, "bad slab plan arguments");
  // key: geometry + everything of the band plans that shapes the iteration
  std::vector<unsigned char> key;
  auto put = [&](const void* p, size_t n) { key.insert(key.end(), (const unsigned char*)p, (const unsigned char*)p + n); };
  put(row_bounds, sizeof(int) * (sl->world + 1));
  put(&cols, sizeof cols);
  for (int b = 0; b < n_bands; ++b) {
    put(&bands[b].kx, sizeof(int)); put(&bands[b].ky, sizeof(int)); put(&bands[b].n_iter, sizeof(int));
    put(&bands[b].direct, sizeof(int));
    put(bands[b].psf_x, sizeof(float) * bands[b].kx); put(bands[b].psf_y, sizeof(float) * bands[b].ky);
  }
  if (changed) *changed = 0;
  if (key == sl->key && sl->arena) return THZ_OK;
  THZ_CUDA(c, cudaStreamSynchronize(sl->stream));
  sl->bounds.assign(row_bounds, row_bounds + sl->world + 1);
  for (int r = 0; r < sl->world; ++r)
    if (sl->bounds[r + 1] <= sl->bounds[r]) return set_err(c, THZ_EINVAL, "empty row slab");
  sl->cols = cols;
  std::vector<SlabBand> nb((size_t)n_bands);
  size_t cursor = 1024;
  for (int b = 0; b < n_bands; ++b) {
    const int rc = slab_band_geometry(sl, bands[b], cursor, nb[b]);
    if (rc != THZ_OK) return rc;
  }
  int what = 1;
  if (cursor > sl->arena_bytes) {
    slab_disconnect(sl);
    if (sl->arena) THZ_CUDA(c, cudaFree(sl->arena));
    sl->arena = nullptr;
    sl->arena_bytes = 0;
    void* p = nullptr;
    THZ_CUDA(c, cudaMalloc(&p, cursor));
    sl->arena = (unsigned char*)p;
    sl->arena_bytes = cursor;
    what = 2;
  }
  // zero flags, counters and the outer halos; the callers hold every rank idle around a plan change
  THZ_CUDA(c, cudaMemsetAsync(sl->arena, 0, sl->arena_bytes, sl->stream));
  THZ_CUDA(c, cudaStreamSynchronize(sl->stream));
  for (int b = 0; b < n_bands; ++b) {
    const int rc = slab_band_finish_plan(sl, bands[b], nb[b]);
    if (rc != THZ_OK) return rc;
  }
  sl->bands.swap(nb);
  sl->key.swap(key);
  if (changed) *changed = what;
  return THZ_OK;
}

int thz_slab_export(thz_slab* sl, void* handle) {
  if (!sl || !handle) return THZ_EINVAL;
  thz_ctx* c = sl->ctx;
  CHECK_CTX(c);
  if (!sl->arena) return set_err(c, THZ_ESTATE, "thz_slab_plan has not been called");
  static_assert(sizeof(cudaIpcMemHandle_t) == THZ_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  THZ_CUDA(c, cudaIpcGetMemHandle(&h, sl->arena));
  memcpy(handle, &h, sizeof h);
  return THZ_OK;
}

int thz_slab_connect_ipc(thz_slab* sl, const void* handles) {
  if (!sl || !handles) return THZ_EINVAL;
  thz_ctx* c = sl->ctx;
  CHECK_CTX(c);
  slab_disconnect(sl);
  const unsigned char* hb = (const unsigned char*)handles;
  auto open = [&](int r, unsigned char** out) -> int {
    cudaIpcMemHandle_t h;
    memcpy(&h, hb + (size_t)r * THZ_IPC_HANDLE_BYTES, sizeof h);
    void* p = nullptr;
    THZ_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *out = (unsigned char*)p;
    return THZ_OK;
  };
  int rc = THZ_OK;
  if (sl->rank > 0) {
    rc = open(sl->rank - 1, &sl->up);
    sl->up_ipc = rc == THZ_OK;
  }
  if (rc == THZ_OK && sl->rank + 1 < sl->world) {
    rc = open(sl->rank + 1, &sl->down);
    sl->down_ipc = rc == THZ_OK;
  }
  sl->connected = rc == THZ_OK;
  return rc;
}

int thz_slab_connect_local(thz_slab* sl, thz_slab* up, thz_slab* down) {
  if (!sl) return THZ_EINVAL;
  thz_ctx* c = sl->ctx;
  CHECK_CTX(c);
  slab_disconnect(sl);
  if ((sl->rank > 0) != (up != nullptr) || (sl->rank + 1 < sl->world) != (down != nullptr))
    return set_err(c, THZ_EINVAL, "neighbours do not match rank / world");
  for (thz_slab* nb : {up, down}) {
    if (!nb) continue;
    if (!nb->arena || nb->arena_bytes != sl->arena_bytes) return set_err(c, THZ_ESTATE, "neighbour holds a different plan");
    if (nb->ctx->device != c->device) {
      int can = 0;
      THZ_CUDA(c, cudaDeviceCanAccessPeer(&can, c->device, nb->ctx->device));
      if (!can) return set_err(c, THZ_ECUDA, "no peer access between the devices of neighbouring ranks");
      cudaError_t e = cudaDeviceEnablePeerAccess(nb->ctx->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) return cuda_fail(c, e, "cudaDeviceEnablePeerAccess");
    }
  }
  sl->up = up ? up->arena : nullptr;
  sl->down = down ? down->arena : nullptr;
  sl->connected = true;
  return THZ_OK;
}

int thz_slab_set_stream(thz_slab* sl, void* stream) {
  if (!sl) return THZ_EINVAL;
  sl->stream = stream ? (cudaStream_t)stream : sl->ctx->stream;
  return THZ_OK;
}

int thz_slab_rl(thz_slab* sl, const float* d_energy, int64_t bstride, float* d_gain, float* d_deconv) {
  if (!sl) return THZ_EINVAL;
  CHECK_CTX(sl->ctx);
  if (!d_energy || !d_gain) return set_err(sl->ctx, THZ_EINVAL, "null pointer");
  thz_slab* ranks[1] = {sl};
  const float* e[1] = {d_energy};
  float* g[1] = {d_gain};
  float* u[1] = {d_deconv};
  return slab_run(ranks, 1, e, &bstride, g, u);
}

int thz_slab_rl_serial(thz_slab* const* ranks, int n, const float* const* d_energy, const int64_t* bstride,
                       float* const* d_gain, float* const* d_deconv) {
  if (!ranks || n < 1 || !d_energy || !bstride || !d_gain) return THZ_EINVAL;
  for (int r = 0; r < n; ++r) {
    if (!ranks[r] || ranks[r]->ctx->device != ranks[0]->ctx->device)
      return set_err(ranks[0] ? ranks[0]->ctx : nullptr, THZ_EINVAL, "serial emulation needs every rank on one device");
    ranks[r]->stream = ranks[0]->ctx->stream;   // one stream, dependency order: no kernel waits on a later launch
  }
  CHECK_CTX(ranks[0]->ctx);
  return slab_run(ranks, n, d_energy, bstride, d_gain, d_deconv);
}

int thz_slab_status(thz_slab* sl) {
  if (!sl) return THZ_EINVAL;
  thz_ctx* c = sl->ctx;
  CHECK_CTX(c);
  THZ_CUDA(c, cudaStreamSynchronize(sl->stream));
  if (!sl->arena) return THZ_OK;
  int err = 0;
  THZ_CUDA(c, cudaMemcpy(&err, sl->arena, sizeof err, cudaMemcpyDeviceToHost));
  if (err) return set_err(c, THZ_ECUDA, "a halo wait timed out: a neighbouring rank did not deliver its boundary rows");
  if (sl->trace && sl->d_trace && sl->trace_launches > 0) {
    std::vector<unsigned long long> h((size_t)sl->trace_launches * 16);
    THZ_CUDA(c, cudaMemcpy(h.data(), sl->d_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(sl->trace_path.c_str(), "w")) {
      fprintf(f, "launch,start_ns,halo_wait_ns,end_ns,boundary_end_ns,top_t0,top_t1,top_t2,top_t3,top_t4,int_t0,int_t1,int_t2,int_t3,int_t4\n");
      for (int i = 0; i < sl->trace_launches; ++i) {
        const unsigned long long* r = h.data() + 16 * (size_t)i;
        fprintf(f, "%d,%llu,%llu,%llu,%llu", i, r[0] - h[0], r[1], r[2] - h[0], r[3] - h[0]);
        for (int k = 4; k < 9; ++k) fprintf(f, ",%lld", (long long)(r[k] - r[0]));
        for (int k = 10; k < 15; ++k) fprintf(f, ",%lld", (long long)(r[k] - r[0]));
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  return THZ_OK;
}

}  // extern "C"
