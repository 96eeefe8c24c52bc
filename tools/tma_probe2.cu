// Probe 2: (a) 1-D bulk copy without descriptor, (b) descriptor in global memory, (c) libcu++ wrappers.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_bulk1d(const float* src, float* out, int n) {
  __shared__ alignas(128) float buf[1024];
  __shared__ alignas(8) uint64_t mb;
  const uint32_t mbar = smem_u32(&mb);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(n * 4) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(buf)), "l"(src), "r"(n * 4), "r"(mbar) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(mbar), "r"(0) : "memory");
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = buf[i];
}

__global__ void k_gdesc(const CUtensorMap* tmap, float* out, int c0, int c1) {
  __shared__ alignas(128) float tile[32][32];
  __shared__ alignas(8) uint64_t mb;
  const uint32_t mbar = smem_u32(&mb);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(4096) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(mbar) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(mbar), "r"(0) : "memory");
  }
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = (&tile[0][0])[i];
}

__global__ void k_cde(const __grid_constant__ CUtensorMap tensor_map, float* out, int x, int y) {
  __shared__ alignas(128) float smem_buffer[32][32];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) {
    init(&bar, blockDim.x);
    cde::fence_proxy_async_shared_cta();
  }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = (&smem_buffer[0][0])[i];
}

int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  int H = 100, W = 80, pitch = 80;
  std::vector<float> h((size_t)H * pitch);
  for (int r = 0; r < H; ++r) for (int c = 0; c < pitch; ++c) h[r * pitch + c] = r * 1000 + c;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4);
  cudaMalloc(&o, 4096);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  auto enc = (PFN_cuTensorMapEncodeTiled_v12000)p;
  alignas(64) CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H};
  cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode %d\n", (int)r);
  int x = argc > 2 ? atoi(argv[2]) : 0, y = argc > 3 ? atoi(argv[3]) : 0;
  if (variant == 0) k_bulk1d<<<1, 128>>>(d, o, 1024);
  else if (variant == 1) {
    CUtensorMap* dm; cudaMalloc(&dm, sizeof(CUtensorMap)); cudaMemcpy(dm, &map, sizeof map, cudaMemcpyHostToDevice);
    k_gdesc<<<1, 128>>>(dm, o, x, y);
  } else k_cde<<<1, 128>>>(map, o, x, y);
  cudaError_t e = cudaDeviceSynchronize();
  printf("variant %d (%d,%d): %s\n", variant, x, y, cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<float> res(1024);
    cudaMemcpy(res.data(), o, 4096, cudaMemcpyDeviceToHost);
    printf("res[0]=%g res[33]=%g res[1023]=%g\n", res[0], res[33], res[1023]);
  }
  return 0;
}
