// Throughput of FFMA (3-register), FFMA2 (fma.rn.f32x2) and HFMA2.BF16 on sm_100a: GFMA/s per SM-clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
constexpr int ITERS = 4096, CHAINS = 8;

__global__ void k_ffma(float* out, float a, float b) {
  float x[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) x[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
  unsigned long long x[CHAINS], av, bv;
  asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    float f = threadIdx.x * 0.001f + i;
    asm("mov.b64 %0, {%1, %1};" : "=l"(x[i]) : "f"(f));
  }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(av), "l"(bv));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_hfma2(float* out, float a, float b) {
  __nv_bfloat162 x[CHAINS], av = __floats2bfloat162_rn(a, a), bv = __floats2bfloat162_rn(b, b);
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) x[i] = __floats2bfloat162_rn(threadIdx.x * 0.001f + i, 1.0f);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = __hfma2(x[i], av, bv);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += __low2float(x[i]) + __high2float(x[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}
int main() {
  float* out;
  const int grid = 148 * 8, block = 512;
  cudaMalloc(&out, grid * block * 4);
  const double n = (double)grid * block * ITERS * CHAINS;
  float t1 = timeit([&] { k_ffma<<<grid, block>>>(out, 0.999f, 0.001f); });
  float t2 = timeit([&] { k_ffma2<<<grid, block>>>(out, 0.999f, 0.001f); });
  float t3 = timeit([&] { k_hfma2<<<grid, block>>>(out, 0.999f, 0.001f); });
  printf("FFMA   %.3f ms  %.1f TFMA/s\n", t1, n / t1 * 1e-9);
  printf("FFMA2  %.3f ms  %.1f TFMA/s (2 per instr)\n", t2, 2 * n / t2 * 1e-9);
  printf("HFMA2.BF16 %.3f ms  %.1f TFMA/s (2 per instr)\n", t3, 2 * n / t3 * 1e-9);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
