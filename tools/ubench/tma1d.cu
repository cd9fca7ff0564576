// does the bulk-copy engine work at all here?  1-D cp.async.bulk (no tensor map), then a 2-D tensor map encoded through -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k1d(const float *src, float *out) {
  __shared__ alignas(128) float buf[1024];
  __shared__ alignas(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(4096u) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf)), "l"(src), "r"(4096u),
                 "r"(smem_u32(&bar))
                 : "memory");
  }
  unsigned ok = 0;
  for (int t = 0; t < (1 << 20) && !ok; t++)
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = buf[i];
}
__global__ void k2d(const __grid_constant__ CUtensorMap tmap, int x, int y, float *out) {
  __shared__ alignas(128) float buf[16 * 32];
  __shared__ alignas(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((unsigned)sizeof(buf)) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(buf)), "l"(&tmap),
                 "r"(x), "r"(y), "r"(smem_u32(&bar))
                 : "memory");
  }
  unsigned ok = 0;
  for (int t = 0; t < (1 << 20) && !ok; t++)
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) out[i] = buf[i];
}
__global__ void k3d(const __grid_constant__ CUtensorMap tmap, int x, int y, int z, float *out, int nbox) {
  extern __shared__ __align__(128) float dbuf[];
  __shared__ alignas(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((unsigned)(nbox * 4)) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dbuf)),
                 "l"(&tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar))
                 : "memory");
  }
  unsigned ok = 0;
  for (int t = 0; t < (1 << 20) && !ok; t++)
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < nbox; i += blockDim.x) out[i] = dbuf[i];
}
typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int run3d(bool entry_point, int bx, int by, int bz, int x) {
  const int N = 64;
  std::vector<float> h((size_t)N * N * N);
  for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) h[((size_t)z * N + y) * N + x] = x + 100.f * y + 10000.f * z;
  float *d, *out;
  cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int nbox = bx * by * bz;
  cudaMalloc(&out, nbox * 4);
  cuInit(0);
  CUtensorMap map;
  cuuint64_t gd[3] = {N, N, N}, gs[2] = {(cuuint64_t)N * 4, (cuuint64_t)N * N * 4};
  cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz}, es[3] = {1, 1, 1};
  CUresult r;
  if (entry_point) {
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    r = ((Enc)p)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                 CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  const int y = 7, z = 20;
  k3d<<<1, 128, nbox * 4>>>(map, x, y, z, out, nbox);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> o(nbox); cudaMemcpy(o.data(), out, nbox * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int kz = 0; kz < bz; kz++) for (int ky = 0; ky < by; ky++) for (int kx = 0; kx < bx; kx++)
    if (o[(kz * by + ky) * bx + kx] != (x + kx) + 100.f * (y + ky) + 10000.f * (z + kz)) bad++;
  printf("3-D x=%d %s box %dx%dx%d: encode=%d sync=%s mismatches=%d\n", x, entry_point ? "entry-point" : "libcuda", bx, by, bz, (int)r, cudaGetErrorString(e), bad);
  return 0;
}
int main(int argc, char **argv) {
  if (argc > 1 && atoi(argv[1]) >= 2) return run3d(atoi(argv[1]) == 3, argc > 2 ? atoi(argv[2]) : 32, argc > 3 ? atoi(argv[3]) : 16, argc > 4 ? atoi(argv[4]) : 8, argc > 5 ? atoi(argv[5]) : 5);
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int N = 64;
  std::vector<float> h((size_t)N * N);
  for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) h[(size_t)y * N + x] = x + 100.f * y;
  float *d, *out;
  cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 4096);
  if (variant == 0) {
    k1d<<<1, 128>>>(d, out);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(1024); cudaMemcpy(o.data(), out, 4096, cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < 1024; i++) if (o[i] != h[i]) bad++;
    printf("1-D bulk copy: sync=%s mismatches=%d\n", cudaGetErrorString(e), bad);
  } else {
    cuInit(0);
    CUtensorMap map;
    cuuint64_t gd[2] = {N, N}, gs[1] = {(cuuint64_t)N * 4};
    cuuint32_t box[2] = {32, 16}, es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("2-D encode via libcuda: %d\n", (int)r);
    k2d<<<1, 128>>>(map, 4, 8, out);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(512); cudaMemcpy(o.data(), out, 2048, cudaMemcpyDeviceToHost);
    int bad = 0; for (int ky = 0; ky < 16; ky++) for (int kx = 0; kx < 32; kx++) if (o[ky * 32 + kx] != (4 + kx) + 100.f * (8 + ky)) bad++;
    printf("2-D tensor load: sync=%s mismatches=%d\n", cudaGetErrorString(e), bad);
  }
  return 0;
}
