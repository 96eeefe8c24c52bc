// Standalone probe for the TMA tile load used by k_rl_conv (debug aid).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VARIANT>
__global__ void k(const __grid_constant__ CUtensorMap tmap, float* out, int box_rows, int box_cols, int c0, int c1) {
  extern __shared__ __align__(128) unsigned char raw[];
  unsigned char* base = raw + ((128u - (smem_u32(raw) & 127u)) & 127u);
  uint64_t* mb = reinterpret_cast<uint64_t*>(base);
  float* tile = reinterpret_cast<float*>(base + 128);
  const uint32_t mbar = smem_u32(mb);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1));
    if (VARIANT & 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    else asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = box_rows * box_cols * 4;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    if (VARIANT & 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(c1), "r"(mbar) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(c0), "r"(c1), "r"(mbar) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(mbar), "r"(0) : "memory");
  }
  for (int i = threadIdx.x; i < box_rows * box_cols; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  int box_cols = argc > 2 ? atoi(argv[2]) : 84;
  int box_rows = argc > 3 ? atoi(argv[3]) : 70;
  int H = 100, W = 77, pitch = 80;
  std::vector<float> h((size_t)H * pitch);
  for (int r = 0; r < H; ++r) for (int c = 0; c < pitch; ++c) h[r * pitch + c] = r * 1000 + c;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, box_rows * box_cols * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  printf("entry %d q %d p %p\n", (int)e, (int)q, p);
  auto enc = (PFN_cuTensorMapEncodeTiled_v12000)p;
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, (variant & 4) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  size_t smem = box_rows * box_cols * 4 + 512;
  int c0 = -3, c1 = -3;
  #define RUN(V) { cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<V><<<1, 256, smem>>>(map, o, box_rows, box_cols, c0, c1); }
  switch (variant & 3) { case 0: RUN(0); break; case 1: RUN(1); break; case 2: RUN(2); break; default: RUN(3); }
  e = cudaDeviceSynchronize();
  printf("variant %d box %dx%d sync: %s\n", variant, box_rows, box_cols, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<float> res(box_rows * box_cols);
    cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < box_rows; ++r2) for (int c = 0; c < box_cols; ++c) {
      int gr = r2 + c1, gc = c + c0;
      float exp = (gr >= 0 && gr < H && gc >= 0 && gc < W) ? gr * 1000 + gc : 0.f;
      if (res[r2 * box_cols + c] != exp) ++bad;
    }
    printf("mismatches %d\n", bad);
  }
  return 0;
}
