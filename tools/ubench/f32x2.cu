// Micro-benchmark: issue rate of scalar FFMA/FADD vs packed FFMA2/FADD2 (sm_100a), and I2F vs PRMT+FADD byte->float.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void k_ffma(float *out, float a, float b) {
  float x[8];
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++)
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = __fmaf_rn(x[i], a, b);
  float s = 0;
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float *out, float a, float b) {
  float2 x[8];
  const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
  for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  for (int it = 0; it < ITERS; it++)
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = __ffma2_rn(x[i], a2, b2);
  float s = 0;
  for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fadd(float *out, float a, float b) {
  float x[8];
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++)
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = __fadd_rn(x[i], a);
  float s = 0;
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fadd2(float *out, float a, float b) {
  float2 x[8];
  const float2 a2 = make_float2(a, a);
  for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  for (int it = 0; it < ITERS; it++)
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = __fadd2_rn(x[i], a2);
  float s = 0;
  for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// byte -> float: I2F (one XU-pipe instruction) vs PRMT into the mantissa of 2^23 + FADD
__global__ void k_i2f(float *out, unsigned w) {
  float s[4] = {0, 0, 0, 0};
  unsigned v = w + threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) s[i] += (float)((v >> (8 * i)) & 0xffu);
    v = v * 1664525u + 1013904223u;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s[0] + s[1] + s[2] + s[3];
}
__global__ void k_prmt(float *out, unsigned w) {
  float s[4] = {0, 0, 0, 0};
  unsigned v = w + threadIdx.x;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 4; i++) s[i] += __uint_as_float(__byte_perm(v, 0x4b000000u, 0x7440 + i)) - 8388608.0f;
    v = v * 1664525u + 1013904223u;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s[0] + s[1] + s[2] + s[3];
}
template <typename F>
static float timeit(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < 5; i++) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / 5;
}
int main() {
  float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int B = 148 * 8, T = 256;
  const double n = (double)B * T * ITERS * 8;
  float t;
  t = timeit([&] { k_ffma<<<B, T>>>(out, 1.0001f, 0.5f); });  printf("FFMA   %.3f ms  %.1f Gthread-instr/s\n", t, n / t / 1e6);
  t = timeit([&] { k_ffma2<<<B, T>>>(out, 1.0001f, 0.5f); }); printf("FFMA2  %.3f ms  %.1f Gthread-instr/s (x2 flops each)\n", t, n / t / 1e6);
  t = timeit([&] { k_fadd<<<B, T>>>(out, 1.0001f, 0.5f); });  printf("FADD   %.3f ms  %.1f Gthread-instr/s\n", t, n / t / 1e6);
  t = timeit([&] { k_fadd2<<<B, T>>>(out, 1.0001f, 0.5f); }); printf("FADD2  %.3f ms  %.1f Gthread-instr/s (x2 flops each)\n", t, n / t / 1e6);
  const double n4 = (double)B * T * ITERS * 4;
  t = timeit([&] { k_i2f<<<B, T>>>(out, 12345u); });  printf("I2F    %.3f ms  %.1f Gconv/s\n", t, n4 / t / 1e6);
  t = timeit([&] { k_prmt<<<B, T>>>(out, 12345u); }); printf("PRMT   %.3f ms  %.1f Gconv/s\n", t, n4 / t / 1e6);
  return 0;
}
