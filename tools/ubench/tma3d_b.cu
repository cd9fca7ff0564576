// variant study for the 3-D TMA load: argv[1] = 0 official libcu++ wrappers, 1 raw PTX (as gxy_march_tma.cu)
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cde = cuda::device::experimental;
using barrier = cuda::barrier<cuda::thread_scope_block>;
#define BX 32
#define BY 16
#define BZ 8
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k_official(const __grid_constant__ CUtensorMap tmap, int x, int y, int z, float *out) {
  __shared__ alignas(128) float buf[BZ][BY][BX];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_3d_global_to_shared(&buf, &tmap, x, y, z, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(buf));
  } else token = bar.arrive();
  bar.wait(std::move(token));
  float *b = &buf[0][0][0];
  for (int i = threadIdx.x; i < BZ * BY * BX; i += blockDim.x) out[i] = b[i];
}
__global__ void k_raw(const __grid_constant__ CUtensorMap tmap, int x, int y, int z, float *out) {
  __shared__ alignas(128) float buf[BZ * BY * BX];
  __shared__ alignas(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(buf)),
                 "l"(&tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar))
                 : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((unsigned)sizeof(buf)) : "memory");
  }
  unsigned ok = 0;
  for (int t = 0; t < (1 << 20) && !ok; t++)
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < BZ * BY * BX; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0, l2 = argc > 2 ? atoi(argv[2]) : 0;
  const int N = 64;
  std::vector<float> h((size_t)N * N * N);
  for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) h[((size_t)z * N + y) * N + x] = x + 100.f * y + 10000.f * z;
  float *d, *out;
  cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, BX * BY * BZ * 4);
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap map;
  cuuint64_t gd[3] = {N, N, N}, gs[2] = {(cuuint64_t)N * 4, (cuuint64_t)N * N * 4};
  cuuint32_t box[3] = {BX, BY, BZ}, es[3] = {1, 1, 1};
  CUresult r = ((Enc)p)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        (CUtensorMapL2promotion)l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("variant %d l2promo %d encode: %d\n", variant, l2, (int)r);
  const int x = 5, y = 7, z = 20;
  if (variant == 0) k_official<<<1, 128>>>(map, x, y, z, out); else k_raw<<<1, 128>>>(map, x, y, z, out);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> o(BX * BY * BZ);
  cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int kz = 0; kz < BZ; kz++) for (int ky = 0; ky < BY; ky++) for (int kx = 0; kx < BX; kx++)
    if (o[(kz * BY + ky) * BX + kx] != (x + kx) + 100.f * (y + ky) + 10000.f * (z + kz)) bad++;
  printf("  sync=%s mismatches=%d\n", cudaGetErrorString(e), bad);
  return 0;
}
