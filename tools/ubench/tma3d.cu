// Stand-alone check of the 3-D TMA box load used by gxy_march_tma.cu (tensor map in global memory and as a __grid_constant__ parameter).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma3d tma3d.cu ; run on the GPU box.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#define BX 32
#define BY 16
#define BZ 8
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <bool PARAM>
__global__ void k(const __grid_constant__ CUtensorMap pmap, const CUtensorMap *gmap, int x, int y, int z, float *out, int *err) {
  __shared__ __align__(128) float buf[BZ * BY * BX];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the barrier init must be visible to the async proxy (TMA)
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const CUtensorMap *m = PARAM ? &pmap : gmap;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"((unsigned)sizeof(buf)) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(buf)),
                 "l"(m), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar))
                 : "memory");
  }
  unsigned ok = 0;
  for (int t = 0; t < (1 << 20) && !ok; t++)
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
  if (!ok && threadIdx.x == 0) *err = 1;
  for (int i = threadIdx.x; i < BZ * BY * BX; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int N = 64;
  std::vector<float> h((size_t)N * N * N);
  for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) h[((size_t)z * N + y) * N + x] = x + 100.f * y + 10000.f * z;
  float *d, *out; int *err;
  cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, BX * BY * BZ * 4); cudaMalloc(&err, 4); cudaMemset(err, 0, 4);
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap map;
  cuuint64_t gd[3] = {N, N, N}, gs[2] = {(cuuint64_t)N * 4, (cuuint64_t)N * N * 4};
  cuuint32_t box[3] = {BX, BY, BZ}, es[3] = {1, 1, 1};
  CUresult r = ((Enc)p)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d (entry point query %d)\n", (int)r, (int)q);
  CUtensorMap *gmap; cudaMalloc(&gmap, sizeof map); cudaMemcpy(gmap, &map, sizeof map, cudaMemcpyHostToDevice);
  for (int variant = 0; variant < 2; variant++) {
    const int x = 5, y = 7, z = 60;  // z runs out of bounds: zero fill expected for z >= 64
    if (variant == 0) k<false><<<1, 128>>>(map, gmap, x, y, z, out, err); else k<true><<<1, 128>>>(map, gmap, x, y, z, out, err);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> o(BX * BY * BZ); int herr = 0;
    cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int kz = 0; kz < BZ; kz++) for (int ky = 0; ky < BY; ky++) for (int kx = 0; kx < BX; kx++) {
      const float want = (z + kz < N) ? (x + kx) + 100.f * (y + ky) + 10000.f * (z + kz) : 0.f;
      if (o[(kz * BY + ky) * BX + kx] != want) bad++;
    }
    printf("%s: sync=%s timeout=%d mismatches=%d\n", variant ? "param map " : "global map", cudaGetErrorString(e), herr, bad);
    if (e != cudaSuccess) break;
  }
  return 0;
}
