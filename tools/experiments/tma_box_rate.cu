// How fast does one SM's TMA fill shared memory with a haloed conv tile, as a function of the ROW WIDTH of the box?
// Box {C ch, 10 px, 18 rows} of a bf16 NHWC image, C = 64 (128-byte rows, the conv kernel's box; the tensor has only `cin` real
// channels, the rest is out-of-bounds zero fill) against C = 16 / 32 (32- / 64-byte rows).  All 148 SMs load concurrently.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_box_rate tma_box_rate.cu -lcuda && ./tma_box_rate
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(32, 1) k_rate(const __grid_constant__ CUtensorMap tm, int iters, int box_bytes, int W, int H, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  uint8_t* buf = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~uintptr_t(1023));
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int slot_bytes = (box_bytes + 1023) / 1024 * 1024;
  long long t0 = 0;
  if (threadIdx.x == 0) {
    t0 = clock64();
    for (int it = 0; it < iters + 4; ++it) {
      const int s = it & 3;
      if (it >= 4) {                                 // wait for the load issued four iterations ago on this slot
        const uint32_t par = ((it - 4) >> 2) & 1;
        asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar[s])), "r"(par) : "memory");
      }
      if (it < iters) {
        const int x = ((blockIdx.x * 7 + it * 13) % (W - 10)), y = ((blockIdx.x * 3 + it * 5) % (H - 18));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(buf + s * slot_bytes)),
                     "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bar[s])), "r"(0), "r"(x), "r"(y), "r"(0)
                     : "memory");
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const int W = 2040, H = 1356, CS = 64;                    // physical pixel stride: 64 channels
  __nv_bfloat16* img;
  cudaMalloc(&img, (size_t)W * H * CS * 2);
  cudaMemset(img, 0, (size_t)W * H * CS * 2);
  long long* dcy;
  cudaMalloc(&dcy, 148 * 8);
  struct { int cin, boxc; CUtensorMapSwizzle sw; const char* name; } cases[] = {
      {64, 64, CU_TENSOR_MAP_SWIZZLE_128B, "cin 64, box 64 ch (128 B rows, all real)"},
      {16, 64, CU_TENSOR_MAP_SWIZZLE_128B, "cin 16, box 64 ch (128 B rows, 48 ch zero fill)"},
      {32, 32, CU_TENSOR_MAP_SWIZZLE_64B, "cin 32, box 32 ch (64 B rows)"},
      {16, 16, CU_TENSOR_MAP_SWIZZLE_32B, "cin 16, box 16 ch (32 B rows)"},
  };
  for (auto& c : cases) {
    CUtensorMap tm;
    cuuint64_t dims[4] = {(cuuint64_t)c.cin, (cuuint64_t)W, (cuuint64_t)H, 1};
    cuuint64_t strides[3] = {(cuuint64_t)CS * 2, (cuuint64_t)W * CS * 2, (cuuint64_t)W * H * CS * 2};
    cuuint32_t box[4] = {(cuuint32_t)c.boxc, 10, 18, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, img, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    const int box_bytes = c.boxc * 2 * 10 * 18, iters = 400;
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 24 * 1024 + 1024);
    k_rate<<<148, 32, 4 * 24 * 1024 + 1024>>>(tm, iters, box_bytes, W, H, dcy);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    long long h[148];
    cudaMemcpy(h, dcy, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += (double)h[i];
    avg /= 148.0 * iters;
    printf("%-52s %6d B per box: %7.0f cycles per box (4 in flight), %5.1f B/clk/SM\n", c.name, box_bytes, avg, box_bytes / avg);
  }
  return 0;
}
