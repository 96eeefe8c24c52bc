// thz_chain_fused.cu -- launcher of the fused trace + band-energy kernel (k_chain_energy_fused, thz_deconv_dev.cuh).
// A translation unit of its own: its 30 instantiations (5 trace lengths x 3 gate modes x 2 hand-off forms) are the
// longest compile of the library and build beside thz_deconv.cu instead of after it.
#include "thz_deconv_dev.cuh"

namespace thz {

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

// ---- trace pass + band energies in one cube pass (k_chain_energy_fused) ----
template <int N, int POST, bool SPEC>
static int launch_chain_fused(thz_ctx* c, cudaStream_t s, const TraceArgs& ta, const FirArgs& fa) {
  using GEO = Geo<N>;
  auto kernel = k_chain_energy_fused<N, POST, SPEC>;
  const size_t smem = FGeo<N>::smem_bytes;
  const void* key = (const void*)kernel;
  auto it = c->occ.find(key);
  if (it == c->occ.end()) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaFuncSetAttribute(smem)");
    // shared memory for kMinBlocks CTAs (exchange buffer + stash, 1 KB reserved per CTA), the rest stays L1 for the tables
    int pct = (int)((GEO::kMinBlocks * (smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024)) + 2;
    if (pct > 100) pct = 100;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaFuncSetAttribute(carveout)");
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, GEO::NT, smem);
    if (e != cudaSuccess) return cuda_fail(c, e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (nb < 1) return set_err(c, THZ_ECUDA, "fused chain kernel does not fit on an SM");
    it = c->occ.emplace(key, nb).first;
  }
  const int64_t npairs = (ta.P + 1) / 2;
  const int64_t nitems = (npairs + GEO::G - 1) / GEO::G;
  if (nitems <= 0) return THZ_OK;
  int64_t grid = (int64_t)c->sm_count * it->second;
  if (grid > nitems) grid = nitems;
  kernel<<<(unsigned)grid, GEO::NT, smem, s>>>(ta, fa);
  c->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(c, e, "fused chain kernel launch");
  return THZ_OK;
}
template <int N> static int do_chain_fused(thz_ctx* c, cudaStream_t s, const TraceArgs& ta, const FirArgs& fa, int post) {
  if constexpr (N < 512) {
    return THZ_EINVAL;
  } else {
    if (fa.edges != nullptr) {
      if (post == 0) return launch_chain_fused<N, 0, true>(c, s, ta, fa);
      if (post == 1) return launch_chain_fused<N, 1, true>(c, s, ta, fa);
      return launch_chain_fused<N, 2, true>(c, s, ta, fa);
    }
    if (post == 0) return launch_chain_fused<N, 0, false>(c, s, ta, fa);
    if (post == 1) return launch_chain_fused<N, 1, false>(c, s, ta, fa);
    return launch_chain_fused<N, 2, false>(c, s, ta, fa);
  }
}
int dispatch_chain_fused(thz_ctx* c, cudaStream_t s, int n, const TraceArgs& ta, const FirArgs& fa, int post) {
  THZ_DISPATCH_M(n, do_chain_fused, c, s, ta, fa, post);
}

}  // namespace thz
