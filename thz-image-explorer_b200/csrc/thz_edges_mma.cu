// thz_edges_mma.cu -- pass-A edge energies on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a.
//
// The band energies of `Deconvolution::filter` (src/filters/deconvolution.rs:963-966) are computed as the Parseval
// energy of the FULL linear convolution minus the energy of the 2 x 249 samples that `convolve1d` cuts off
// (:279, 311-315).  Those cut-off samples are
//     head_b[k] = sum_{j <= k} h_b[k - j] x[j]              k, j in [0, 249)   (first 249 taps, first 249 samples)
//     tail_b[k] = sum_{j >= k} h_b[j - k] x[N - 249 + j]                       (taps are symmetric)
// i.e. for every band and edge ONE fixed 249 x 249 triangular Toeplitz matrix applied to every trace of the cube:
// the one true dense contraction on the path.  Per edge  D[P x (8 x 256)] = X_e[P x 256] . T^T  with
// K = 256, followed by a row-wise sum of squares per band.  On CUDA cores this costs nine 512-point transforms per
// trace pair and edge (k_fir_edges); here it runs as tcgen05.mma kind::tf32 with the accumulators in tensor memory.
//
// Precision: both operands are rounded to TF32 with round-to-nearest (the hardware would truncate raw fp32, which
// biases the energies by -5e-4); the products are exact in the fp32 accumulator.  The edge energies are a
// correction of a few per cent of a band energy, so that the measured error of the band energies stays below
// 1e-5 for n >= 2048 (tests/test_edges_mma_gpu.py); shorter traces keep the transform kernel.
//
// One persistent CTA per SM, 8 warps:
//   warp 0      B producer: 32 KB Toeplitz tiles (one K block of one band and edge, stored in global memory as
//               the exact shared-memory image) by cp.async.bulk into a 3-stage ring
//   warp 1      MMA issuer (one elected thread) and owner of the 512 TMEM columns (two 128 x 256 accumulators)
//   warps 2, 3  A loaders: 128 traces x 32 samples per K block from the cube, rounded to TF32, stored in the
//               128-byte-swizzled K-major layout; every K block is reused by the 8 bands
//   warps 4..7  epilogue: tcgen05.ld of the accumulator rows (one trace per thread), sum of squares per band,
//               subtraction from the band energies
// The matrices are triangular, but skipping their all-zero quarter (N = 128 MMAs on 16 KB half tiles, 12 of 16
// tiles per band and edge) measured SLOWER than the dense N = 256 form (14.1 vs 11.4 ms at config 5): twice as many
// mbarrier round trips on the single issuing thread cost more than the 25 % of tensor work they save.
#include "thz_internal.h"

#include <cuda.h>
#include <math.h>
#include <algorithm>
#include <vector>

namespace thz {

namespace {

constexpr int kEM = 128;                   // traces per tile = UMMA M
constexpr int kEN = 256;                   // outputs per band and edge (249 used) = UMMA N
constexpr int kEK = 256;                   // samples per edge (249 used)
constexpr int kKB = 32;                    // K block: one 128-byte swizzle row of tf32
constexpr int kNKB = kEK / kKB;            // 8
constexpr int kBStages = 3;
constexpr int kSeg = (THZ_FIR_TAPS - 1) / 2;             // 249
constexpr uint32_t kATile = kEM * kKB * 4;               // 16 KB
constexpr uint32_t kBTile = kEN * kKB * 4;               // 32 KB
constexpr uint32_t kSmemA = kATile * kNKB;               // 128 KB
constexpr uint32_t kSmemB = kBTile * kBStages;           // 96 KB
constexpr uint32_t kSmemBar = 256;
constexpr uint32_t kSmemTotal = kSmemA + kSmemB + kSmemBar + 1024;   // + alignment slack
constexpr int kThreads = 256;
constexpr int kMaxBandsMma = 8;            // bands per launch (register array of the epilogue)

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// shared-memory matrix descriptor, K-major, 128-byte swizzle (8 rows x 128 B atoms, 1024 B apart)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address
  d |= (uint64_t)1 << 16;                            // leading byte offset: unused for swizzled K-major (canonical value 1)
  d |= (uint64_t)(1024u >> 4) << 32;                 // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 256
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kEN >> 3) << 17) | ((uint32_t)(kEM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

struct EdgeMmaArgs {
  const float* x;          // [P][n] filtered cube
  int n;
  int64_t P;
  int B;                   // bands of this launch (<= kMaxBandsMma)
  const float* tmat;       // [2 edges][B][kNKB] tiles of kBTile bytes: Toeplitz blocks as shared-memory images
  float* energy;           // [B][bstride] full-convolution energies in, "same"-window energies out
  int64_t bstride;
};

__global__ void __launch_bounds__(kThreads, 1) k_fir_edges_mma(const EdgeMmaArgs a) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = smem_raw + ((1024u - (s_addr(smem_raw) & 1023u)) & 1023u);
  unsigned char* sA = base;
  unsigned char* sB = base + kSmemA;
  const uint32_t bars = s_addr(base + kSmemA + kSmemB);
  // barrier map (8 bytes each)
  const uint32_t bar_a_full = bars, bar_a_empty = bars + 8 * kNKB;
  const uint32_t bar_b_full = bars + 16 * kNKB, bar_b_empty = bar_b_full + 8 * kBStages;
  const uint32_t bar_t_full = bar_b_empty + 8 * kBStages, bar_t_empty = bar_t_full + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + kSmemA + kSmemB + 240);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ntiles = (a.P + kEM - 1) / kEM;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kNKB; ++i) {
      bar_init(bar_a_full + 8 * i, 64);     // every loader thread arrives
      bar_init(bar_a_empty + 8 * i, 1);     // tcgen05.commit
    }
    for (int i = 0; i < kBStages; ++i) {
      bar_init(bar_b_full + 8 * i, 1);
      bar_init(bar_b_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      bar_init(bar_t_full + 8 * i, 1);
      bar_init(bar_t_empty + 8 * i, 128);   // every epilogue thread arrives
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // 512 TMEM columns: two 128 x 256 fp32 accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_addr(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- B producer
    if (lane == 0) {
      uint32_t use = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int edge = 0; edge < 2; ++edge)
          for (int b = 0; b < a.B; ++b)
            for (int kb = 0; kb < kNKB; ++kb, ++use) {
              const uint32_t st = use % kBStages, round = use / kBStages;
              bar_wait(bar_b_empty + 8 * st, (round & 1) ^ 1);
              bar_expect_tx(bar_b_full + 8 * st, kBTile);
              const float* src = a.tmat + ((size_t)((edge * a.B + b) * kNKB + kb)) * (kBTile / 4);
              bulk_g2s(s_addr(sB + st * kBTile), src, kBTile, bar_b_full + 8 * st);
            }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      uint32_t buse = 0, tuse = 0, ause = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
        for (int edge = 0; edge < 2; ++edge, ++ause)
          for (int b = 0; b < a.B; ++b, ++tuse) {
            const uint32_t buf = tuse & 1;
            bar_wait(bar_t_empty + 8 * buf, ((tuse >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_tmem = tmem_base + buf * kEN;
            for (int kb = 0; kb < kNKB; ++kb, ++buse) {
              if (b == 0) bar_wait(bar_a_full + 8 * kb, ause & 1);      // K block of this edge has landed
              const uint32_t st = buse % kBStages;
              bar_wait(bar_b_full + 8 * st, (buse / kBStages) & 1);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint64_t da = umma_desc(s_addr(sA + kb * kATile));
              const uint64_t db = umma_desc(s_addr(sB + st * kBTile));
#pragma unroll
              for (int k = 0; k < kKB / 8; ++k)                          // UMMA K = 8 tf32 = 32 bytes
                umma_tf32(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), (kb | k) != 0);
              umma_commit(bar_b_empty + 8 * st);                         // tile free once these MMAs are done
              if (b == a.B - 1) umma_commit(bar_a_empty + 8 * kb);       // last band: the K block may be refilled
            }
            umma_commit(bar_t_full + 8 * buf);
          }
    }
  } else if (warp < 4) {
    // ---------------------------------------------------------------- A loaders (64 threads)
    const int lt = threadIdx.x - 64;
    uint32_t ause = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
      for (int edge = 0; edge < 2; ++edge, ++ause) {
        const int64_t p0 = tile * kEM;
        const int col0 = edge ? a.n - kEK : 0;
        for (int kb = 0; kb < kNKB; ++kb) {
          bar_wait(bar_a_empty + 8 * kb, (ause & 1) ^ 1);
          // 128 rows x 8 float4: thread -> (row = i * 8 + lt / 8, 16-byte chunk = lt % 8), i < 16
          float4 v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int row = i * 8 + (lt >> 3), ch = lt & 7;
            const int64_t p = p0 + row;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < a.P) v[i] = __ldg(reinterpret_cast<const float4*>(a.x + p * a.n + col0 + kb * kKB) + ch);
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int row = i * 8 + (lt >> 3), ch = lt & 7;
            const float4 t = make_float4(to_tf32(v[i].x), to_tf32(v[i].y), to_tf32(v[i].z), to_tf32(v[i].w));
            unsigned char* dst = sA + kb * kATile + (row >> 3) * 1024 + (row & 7) * 128 + ((ch ^ (row & 7)) << 4);
            *reinterpret_cast<float4*>(dst) = t;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> tensor-core reads
          bar_arrive(bar_a_full + 8 * kb);
        }
      }
  } else {
    // ---------------------------------------------------------------- epilogue (warps 4..7 <-> TMEM lanes 0..127)
    const int row = (warp & 3) * 32 + lane;
    uint32_t tuse = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      float e[kMaxBandsMma];
#pragma unroll
      for (int b = 0; b < kMaxBandsMma; ++b) e[b] = 0.f;
      for (int edge = 0; edge < 2; ++edge)
#pragma unroll
        for (int b = 0; b < kMaxBandsMma; ++b) {
          if (b < a.B) {
            const uint32_t buf = tuse & 1;
            bar_wait(bar_t_full + 8 * buf, (tuse >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + buf * kEN + ((uint32_t)((warp & 3) * 32) << 16);
            float s = 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < kEN; c0 += 32) {
              uint32_t r[32];
              asm volatile(
                  "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                  "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                  "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                  : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                    "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                    "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                    "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                  : "r"(taddr + (uint32_t)c0)
                  : "memory");
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float v = __uint_as_float(r[i]);
                s = fmaf(v, v, s);
              }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive(bar_t_empty + 8 * buf);
            e[b] += s;
            ++tuse;
          }
        }
      const int64_t p = tile * kEM + row;
      if (p < a.P) {
#pragma unroll
        for (int b = 0; b < kMaxBandsMma; ++b)
          if (b < a.B) {
            float* dst = a.energy + (size_t)b * a.bstride + p;
            const float cur = *dst;
            if (cur != 0.f) *dst = fmaxf(cur - e[b], 0.f);   // exact zeros (dead pixels) stay zero
          }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

float host_tf32(float x) {   // round to nearest, ties away (cvt.rna), like the device side
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  float y;
  memcpy(&y, &u, 4);
  return y;
}

}  // namespace

// Toeplitz blocks of all bands and both edges as shared-memory images: tile (edge, band, kb) holds rows n = output
// sample, 32 columns j = input sample 32 kb .. 32 kb + 31, 128-byte swizzled like the A tiles.
//   head:  T[n][j] = h[n - j]          j <= n < 249                       (input column j = sample j)
//   tail:  T[n][j] = h[498 - ((j - 7) - n)]   n <= j - 7 < 249            (input column j = sample N - 256 + j)
static void build_edge_matrices(const thz_band_plan* bands, int B, std::vector<float>& out) {
  out.assign((size_t)2 * B * kNKB * (kBTile / 4), 0.f);
  for (int edge = 0; edge < 2; ++edge)
    for (int b = 0; b < B; ++b) {
      const float* h = bands[b].fir;
      for (int kb = 0; kb < kNKB; ++kb) {
        float* tile = out.data() + ((size_t)((edge * B + b) * kNKB + kb)) * (kBTile / 4);
        for (int n = 0; n < kEN; ++n)
          for (int jj = 0; jj < kKB; ++jj) {
            const int j = kb * kKB + jj;
            float v = 0.f;
            if (n < kSeg) {
              if (edge == 0) {
                if (j <= n) v = h[n - j];
              } else {
                const int js = j - (kEK - kSeg);   // sample index inside the last 249
                if (js >= n && js < kSeg) v = h[THZ_FIR_TAPS - 1 - (js - n)];   // taps 250 .. 498 (= h[js - n] by symmetry)
              }
            }
            const int ch = jj >> 2, w = jj & 3;
            const size_t off = (size_t)(n >> 3) * 256 + (size_t)(n & 7) * 32 + (size_t)((ch ^ (n & 7)) << 2) + w;
            tile[off] = host_tf32(v);
          }
      }
    }
}

bool edges_mma_supported(int n) { return n >= 2048 && (n % 4) == 0; }

// Subtracts the head / tail energies from d_energy for all bands (chunks of <= 8 bands per launch).
// edge_rows: d_cube is the side buffer of the spectral hand-off, rows of 512 floats = first / last 256 samples of
// every trace (n = 512: the kernel only reads columns [0, 256) and [n - 256, n))
int launch_fir_edges_mma(thz_ctx* c, cudaStream_t s, const float* d_cube, int64_t P, int n, const thz_band_plan* bands,
                         int B, float* d_energy, int64_t bstride, bool edge_rows) {
  if (edge_rows ? n != 512 : !edges_mma_supported(n))
    return set_err(c, THZ_EINVAL, "tensor-core edge pass needs n >= 2048");
  // matrices are cached per context, keyed by the taps
  uint64_t key = 0xcbf29ce484222325ull;
  for (int b = 0; b < B; ++b)
    for (size_t i = 0; i < sizeof(bands[b].fir); ++i) {
      key ^= reinterpret_cast<const unsigned char*>(bands[b].fir)[i];
      key *= 0x100000001B3ull;
    }
  const size_t bytes = (size_t)2 * B * kNKB * kBTile;
  void* dm = nullptr;
  int rc = ws_get(c, WS_EDGE_MMA, bytes, &dm);
  if (rc != THZ_OK) return rc;
  if (c->edge_mma_key != key || c->edge_mma_bands != B) {
    std::vector<float> host;
    // tiles are laid out per launch group of <= 8 bands: [group][edge][band in group][kb]
    std::vector<float> all;
    for (int b0 = 0; b0 < B; b0 += kMaxBandsMma) {
      const int nb = std::min(kMaxBandsMma, B - b0);
      build_edge_matrices(bands + b0, nb, host);
      all.insert(all.end(), host.begin(), host.end());
    }
    THZ_CUDA(c, cudaStreamSynchronize(s));
    THZ_CUDA(c, cudaMemcpyAsync(dm, all.data(), all.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    THZ_CUDA(c, cudaStreamSynchronize(s));
    c->edge_mma_key = key;
    c->edge_mma_bands = B;
  }
  cudaError_t e = ensure_dynamic_smem(c, (const void*)k_fir_edges_mma, kSmemTotal);
  if (e != cudaSuccess) return cuda_fail(c, e, "cudaFuncSetAttribute(k_fir_edges_mma)");
  const int64_t ntiles = (P + kEM - 1) / kEM;
  const int grid = (int)std::min<int64_t>(ntiles, c->sm_count);
  size_t off = 0;
  for (int b0 = 0; b0 < B; b0 += kMaxBandsMma) {
    const int nb = std::min(kMaxBandsMma, B - b0);
    EdgeMmaArgs a{};
    a.x = d_cube; a.n = n; a.P = P; a.B = nb;
    a.tmat = (const float*)dm + off;
    a.energy = d_energy + (size_t)b0 * bstride;
    a.bstride = bstride;
    k_fir_edges_mma<<<grid, kThreads, kSmemTotal, s>>>(a);
    c->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(c, e, "k_fir_edges_mma launch");
    off += (size_t)2 * nb * kNKB * (kBTile / 4);
  }
  return THZ_OK;
}

}  // namespace thz
