// Micro-experiment 2: as umma_noswizzle.cu, but the B operand is MN-MAJOR (idesc bit 16): B[k][n] stored as planes
//   plane[n / 8][k] = 16 bytes = 8 consecutive n of one k  (the layout of attention K/V rows gathered as 8-channel cells),
// descriptor LBO = 128 B (next 8 k), SBO = plane stride (next 8 n); variants try the swapped roles too.
// Micro-experiment: tcgen05.mma with SWIZZLE_NONE (interleaved core-matrix) K-major operands whose start address is an
// arbitrary 16-byte multiple.  A lives in "8-channel planes": plane[kg][pixel] = 16 bytes (8 bf16 channels of one pixel);
// an M=128 operand = 16 groups of 8 consecutive pixels, groups SBO bytes apart, the two 8-channel halves of a K=16 step
// LBO bytes apart.  Prints max |D - reference| for a few (start shift, SBO) combinations.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_noswizzle umma_noswizzle.cu && ./umma_noswizzle
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int NPIX = 34 * 34;      // haloed buffer: 34 x 34 pixels
constexpr int PW = 34;
constexpr int KG = 2;              // 16 channels = one K step
constexpr int NOUT = 32;
constexpr int KROWS = 64;          // k rows held per plane (the MMA reads 16 of them from row k0)

__global__ void __launch_bounds__(128, 1) k_test(const __nv_bfloat16* __restrict__ act /*[KG][NPIX][8]*/,
                                                 const __nv_bfloat16* __restrict__ wgt /*[NOUT/8][KROWS][8]*/, float* __restrict__ out /*[128][NOUT]*/,
                                                 int start_pix, int sbo_bytes, int b_lbo, int b_sbo, int k0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* sA = smem;                                  // KG planes x NPIX x 16 B
  uint8_t* sB = smem + KG * NPIX * 16;                 // NOUT/8 planes x KROWS x 16 B
  for (int i = threadIdx.x; i < KG * NPIX * 8; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(sA)[i] = act[i];
  for (int i = threadIdx.x; i < (NOUT / 8) * KROWS * 8; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(sB)[i] = wgt[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> async proxy (tensor core)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    // descriptors: start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | layout 0 (no swizzle)
    const uint32_t a_addr = smem_u32(sA) + (uint32_t)start_pix * 16u;
    const uint64_t da = (uint64_t)((a_addr & 0x3FFFF) >> 4) | ((uint64_t)((NPIX * 16) >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
    const uint32_t b_addr = smem_u32(sB) + (uint32_t)k0 * 16u;
    const uint64_t db = (uint64_t)((b_addr & 0x3FFFF) >> 4) | ((uint64_t)(b_lbo >> 4) << 16) | ((uint64_t)(b_sbo >> 4) << 32) | (1ull << 46);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(NOUT >> 3) << 17) | ((128u >> 4) << 24);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 1;\n\tsetp.eq.b32 p, 0, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc) : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everyone waits for the MMA
  asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[32];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                 "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                 "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int k = 0; k < 32; ++k) out[(warp * 32 + lane) * NOUT + k] = __uint_as_float(v[k]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

int main() {
  const int na = KG * NPIX * 8, nw = (NOUT / 8) * KROWS * 8;
  __nv_bfloat16 *ha = (__nv_bfloat16*)malloc(na * 2), *hw = (__nv_bfloat16*)malloc(nw * 2);
  float *fa = (float*)malloc(na * 4), *fw = (float*)malloc(nw * 4);
  srand(1);
  for (int i = 0; i < na; ++i) { ha[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fa[i] = __bfloat162float(ha[i]); }
  for (int i = 0; i < nw; ++i) { hw[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fw[i] = __bfloat162float(hw[i]); }
  __nv_bfloat16 *da, *dw; float* dout;
  cudaMalloc(&da, na * 2); cudaMalloc(&dw, nw * 2); cudaMalloc(&dout, 128 * NOUT * 4);
  cudaMemcpy(da, ha, na * 2, cudaMemcpyHostToDevice); cudaMemcpy(dw, hw, nw * 2, cudaMemcpyHostToDevice);
  const int smem = KG * NPIX * 16 + (NOUT / 8) * KROWS * 16;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  float* hout = (float*)malloc(128 * NOUT * 4);
  const int variants[][3] = {{128, KROWS * 16, 0}, {128, KROWS * 16, 16}, {KROWS * 16, 128, 0}, {KROWS * 16, 128, 16}};   // {LBO, SBO, first k row}
  for (int vi = 0; vi < 4; ++vi) {
    const int sp = 35, b_lbo = variants[vi][0], b_sbo = variants[vi][1], k0 = variants[vi][2];
    cudaMemset(dout, 0, 128 * NOUT * 4);
    k_test<<<1, 128, smem>>>(da, dw, dout, sp, PW * 16, b_lbo, b_sbo, k0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", vi, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hout, dout, 128 * NOUT * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int m = 0; m < 128; ++m) {
      const int pix = sp + (m / 8) * PW + (m % 8);
      for (int n = 0; n < NOUT; ++n) {
        double acc = 0;
        for (int k = 0; k < 16; ++k) acc += (double)fa[((k / 8) * NPIX + pix) * 8 + k % 8] * fw[((n / 8) * KROWS + k0 + k) * 8 + n % 8];
        worst = fmax(worst, fabs(acc - hout[m * NOUT + n]));
      }
    }
    printf("B MN-major, LBO %5d SBO %5d k0 %2d: max |D - ref| = %.3e\n", b_lbo, b_sbo, k0, worst);
  }
  return 0;
}
