// Phase 4 token pipeline (bf16 mode) as two tile-resident tcgen05 kernels; tokens never leave the SM between the layers.
//   k_token_attn : t1 = x + out_proj(MHA(LN1(x)))      over the 4 expert tokens of each LR pixel   large_kernel_attention.py:389-391
//   k_token_ffn  : t2 = t1 + ffn2(GELU(ffn0(LN2(t1))))                                              large_kernel_attention.py:392
// They replace seven launches (LayerNorm, qkv 1x1, 4-token attention, out 1x1 + residual, LayerNorm, ffn0 1x1 + GELU,
// ffn2 1x1 + residual) that each streamed a [4][H][W][128..384] tensor through HBM.  The four weight matrices (256 KB of bf16)
// do not fit one SM's shared memory together, hence two kernels with 128 KB of resident weights each.
//
// LayerNorm is folded into the contraction that follows it: with g = gamma o W (rounded to bf16 on the host),
//   LN(x) W^T = rstd * (x g^T - mean * colsum(g)) + (W beta + b)
// so the A operand is the RAW bf16 token row exactly as it sits in HBM, and mean / rstd are applied to the fp32 accumulator in
// the epilogue (no bf16 rounding of the normalised row).  Row statistics: each thread owns a quarter row (32 channels), computes
// its local mean and centred sum of squares, and the four quarters are combined with Chan's formula (no E[x^2] - mean^2).
//
// Layout: 512 threads = 16 warps; warp w owns TMEM lanes 32 (w % 4) .. = tile rows, and the channel quarter w / 4.
//   attention kernel: a tile is 32 pixels x 4 experts, row = expert * 32 + pixel, so the four tokens of a pixel sit in the same
//   lane of the four lane groups; K and V cross between the groups through a 32 KB shared exchange buffer in the same plane
//   layout as the MMA operands (16-byte cells, consecutive rows = consecutive cells: conflict-free).
//   operands: no-swizzle K-major planes (tc_ptx.cuh): activations [kg 16][row 128] x 16 B, weights [kg][n] x 16 B.
// The next tile's rows are prefetched into registers while the current tile computes.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/ffsr_b200.h"

namespace {
using namespace tcx;

constexpr int TK_THREADS = 512;
constexpr int TK_PLANE = 2048;                    // one 8-channel plane of a 128-row tile
constexpr int TK_SA = 16 * TK_PLANE;              // 32,768 B

// ---- attention kernel
constexpr int TA_WQKV = 384 * 128 * 2;            // 98,304 B  [kg 16][n 384][8]   (q rows pre-scaled by 1/4)
constexpr int TA_WO = 128 * 128 * 2;              // 32,768 B  [kg 16][n 128][8]
constexpr int TA_WBYTES = TA_WQKV + TA_WO;
constexpr int TA_PF = 384 + 384 + 128;            // colsum(g)[384] | folded bias[384] | out_proj bias[128]
constexpr int TA_S_PAR = 64;
constexpr int TA_S_STAT = 4096;                   // float2 [4][128]
constexpr int TA_S_W = 8192;
constexpr int TA_S_A = TA_S_W + TA_WBYTES;
constexpr int TA_S_X = TA_S_A + TK_SA;
constexpr int TA_SMEM = TA_S_X + TK_SA;           // 204,800 B

// ---- FFN kernel
constexpr int TF_W0 = 256 * 128 * 2;              // 65,536 B  [kg 16][n 256][8]
constexpr int TF_W2 = 128 * 256 * 2;              // 65,536 B  [kg 32][n 128][8]
constexpr int TF_WBYTES = TF_W0 + TF_W2;
constexpr int TF_PF = 256 + 256 + 128;            // colsum(g)[256] | folded bias[256] | ffn2 bias[128]
constexpr int TF_S_PAR = 64;
constexpr int TF_S_STAT = 4096;
constexpr int TF_S_W = 8192;
constexpr int TF_S_A = TF_S_W + TF_WBYTES;
constexpr int TF_S_H = TF_S_A + TK_SA;
constexpr int TF_SMEM = TF_S_H + TK_SA;           // 204,800 B

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__device__ __forceinline__ void unpack8(const uint4& c, float* v) {
  v[0] = bf_lo(c.x); v[1] = bf_hi(c.x); v[2] = bf_lo(c.y); v[3] = bf_hi(c.y);
  v[4] = bf_lo(c.z); v[5] = bf_hi(c.z); v[6] = bf_lo(c.w); v[7] = bf_hi(c.w);
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// this thread's quarter row (4 cells) -> operand planes; local mean and centred sum of squares -> stat[cq][row]
__device__ __forceinline__ void put_cells_stats(const uint4 (&xc)[4], uint8_t* sA, int cq, int row, float2* stat) {
  float v[32];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    *reinterpret_cast<uint4*>(sA + (cq * 4 + g) * TK_PLANE + row * 16) = xc[g];
    unpack8(xc[g], v + 8 * g);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += v[i];
  const float m = s * (1.0f / 32.0f);
  float m2 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float d = v[i] - m;
    m2 = fmaf(d, d, m2);
  }
  stat[cq * 128 + row] = make_float2(m, m2);
}
// LayerNorm(128) statistics of a row from its four quarter records
__device__ __forceinline__ void row_stats(const float2* stat, int row, float& mean, float& rstd) {
  const float2 a = stat[row], b = stat[128 + row], c = stat[256 + row], d = stat[384 + row];
  mean = ((a.x + b.x) + (c.x + d.x)) * 0.25f;
  const float da = a.x - mean, db = b.x - mean, dc = c.x - mean, dd = d.x - mean;
  const float m2 = ((a.y + b.y) + (c.y + d.y)) + 32.0f * ((da * da + db * db) + (dc * dc + dd * dd));
  rstd = rsqrtf(m2 * (1.0f / 128.0f) + 1e-5f);
}
// 16 accumulator columns -> rstd * (acc - mean * cs) + b      (cs, b: shared memory, 16-byte aligned, warp-uniform address)
__device__ __forceinline__ void fold16(const uint32_t (&v)[16], const float* cs, const float* b, float rstd, float nmr, float* o) {
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    const float4 c = *reinterpret_cast<const float4*>(cs + 4 * q4), bb = *reinterpret_cast<const float4*>(b + 4 * q4);
    o[4 * q4 + 0] = fmaf(rstd, __uint_as_float(v[4 * q4 + 0]), fmaf(nmr, c.x, bb.x));
    o[4 * q4 + 1] = fmaf(rstd, __uint_as_float(v[4 * q4 + 1]), fmaf(nmr, c.y, bb.y));
    o[4 * q4 + 2] = fmaf(rstd, __uint_as_float(v[4 * q4 + 2]), fmaf(nmr, c.z, bb.z));
    o[4 * q4 + 3] = fmaf(rstd, __uint_as_float(v[4 * q4 + 3]), fmaf(nmr, c.w, bb.w));
  }
}
// K = 128 contraction of the activation planes at a32 with weight planes [kg][ntot] starting at row n0, N = n columns
__device__ __forceinline__ void gemm_k128(uint32_t tmem_d, uint32_t a32, uint32_t w32, int ntot, int n0, int n, bool overwrite) {
  const uint32_t hi = desc_hi(128);
  const uint32_t id = idesc_bf16_m128(n);
  const uint32_t b32 = w32 + (uint32_t)n0 * 16u, lbo = (uint32_t)ntot * 16u;
  umma_taps_1x4(tmem_d, desc_lo(a32, TK_PLANE), desc_lo(b32, lbo), id, overwrite ? 0u : 1u, hi, hi, (2u * TK_PLANE) >> 4, 2u * (uint32_t)ntot, 0u);
  umma_taps_1x4(tmem_d, desc_lo(a32 + 8u * TK_PLANE, TK_PLANE), desc_lo(b32 + 8u * lbo, lbo), id, 1u, hi, hi, (2u * TK_PLANE) >> 4,
                2u * (uint32_t)ntot, 0u);
}

__global__ void __launch_bounds__(TK_THREADS, 1) k_token_attn(const __nv_bfloat16* __restrict__ xin, int B, int HW,
                                                              const uint8_t* __restrict__ wblob, const float* __restrict__ pblob,
                                                              __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8);
  float* par = reinterpret_cast<float*>(smem + TA_S_PAR);
  float2* stat = reinterpret_cast<float2*>(smem + TA_S_STAT);
  uint8_t* sW = smem + TA_S_W;
  uint8_t* sA = smem + TA_S_A;
  uint8_t* sX = smem + TA_S_X;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int e = warp & 3, cq = warp >> 2;
  const int row = e * 32 + lane;

  for (int i = tid; i < TA_WBYTES / 16; i += TK_THREADS) reinterpret_cast<uint4*>(sW)[i] = __ldg(reinterpret_cast<const uint4*>(wblob) + i);
  for (int i = tid; i < TA_PF; i += TK_THREADS) par[i] = __ldg(pblob + i);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float* cs = par;
  const float* bq = par + 384;
  const float* bo = par + 768;
  const uint32_t wq_32 = smem_u32(sW), wo_32 = wq_32 + TA_WQKV, a32 = smem_u32(sA);
  const uint32_t trow = tmem + ((uint32_t)(e * 32) << 16);
  uint32_t phase = 0;
  const int tpi = (HW + 31) / 32;
  const long tiles = (long)B * tpi;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

  auto row_ptr = [&](long tile, bool& live) -> long {
    const int b = (int)(tile / tpi), p = (int)(tile - (long)b * tpi) * 32 + lane;
    live = p < HW;
    return (((long)b * 4 + e) * HW + (live ? p : 0)) * 128 + cq * 32;
  };
  uint4 nx[4];
  bool nlive = false;
  if ((long)blockIdx.x < tiles) {
    const long o = row_ptr(blockIdx.x, nlive);
#pragma unroll
    for (int g = 0; g < 4; ++g) nx[g] = nlive ? __ldg(reinterpret_cast<const uint4*>(xin + o) + g) : zero4;
  }

  for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    uint4 xc[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) xc[g] = nx[g];
    const bool live = nlive;
    bool dummy;
    const long orow = row_ptr(tile, dummy);
    put_cells_stats(xc, sA, cq, row, stat);
    if (tile + gridDim.x < tiles) {                  // next tile's rows: in flight during this tile's compute
      const long o = row_ptr(tile + gridDim.x, nlive);
#pragma unroll
      for (int g = 0; g < 4; ++g) nx[g] = nlive ? __ldg(reinterpret_cast<const uint4*>(xin + o) + g) : zero4;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- qkv = x . g^T  (N = 256 + 128) at TMEM columns [0, 384)
    if (warp == 0) {
      tc_fence_after();
      gemm_k128(tmem, a32, wq_32, 384, 0, 256, true);
      gemm_k128(tmem + 256, a32, wq_32, 384, 256, 128, true);
      umma_commit(bar);
    }
    float mean, rstd;
    row_stats(stat, row, mean, rstd);
    const float nmr = -mean * rstd;
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- K of this quarter row -> exchange buffer
    {
      uint32_t v0[16], v1[16];
      tmem_ld16(trow + (uint32_t)(128 + cq * 32), v0);
      tmem_ld16(trow + (uint32_t)(128 + cq * 32 + 16), v1);
      tmem_wait_ld(v0);
      tmem_wait_ld(v1);
      float k[32];
      fold16(v0, cs + 128 + cq * 32, bq + 128 + cq * 32, rstd, nmr, k);
      fold16(v1, cs + 128 + cq * 32 + 16, bq + 128 + cq * 32 + 16, rstd, nmr, k + 16);
#pragma unroll
      for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(sX + (cq * 4 + g) * TK_PLANE + row * 16) = pack8(k + 8 * g);
    }
    __syncthreads();
    // ---- scores and softmax of this row's two heads (q pre-scaled by 1/4 through the weights)
    float pr[2][4];
#pragma unroll
    for (int hd = 0; hd < 2; ++hd) {
      uint32_t v[16];
      tmem_ld16(trow + (uint32_t)(cq * 32 + hd * 16), v);
      tmem_wait_ld(v);
      float q[16];
      fold16(v, cs + cq * 32 + hd * 16, bq + cq * 32 + hd * 16, rstd, nmr, q);
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint8_t* kc = sX + (cq * 4 + hd * 2) * TK_PLANE + (j * 32 + lane) * 16;
        float kk[16];
        unpack8(*reinterpret_cast<const uint4*>(kc), kk);
        unpack8(*reinterpret_cast<const uint4*>(kc + TK_PLANE), kk + 8);
        float a = 0.f;
#pragma unroll
        for (int d = 0; d < 16; ++d) a = fmaf(q[d], kk[d], a);
        pr[hd][j] = a;
        mx = fmaxf(mx, a);
      }
      float den = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        pr[hd][j] = __expf(pr[hd][j] - mx);
        den += pr[hd][j];
      }
      const float inv = __fdividef(1.0f, den);
#pragma unroll
      for (int j = 0; j < 4; ++j) pr[hd][j] *= inv;
    }
    __syncthreads();                                 // every K read is done: the exchange buffer takes V
    {
      uint32_t v0[16], v1[16];
      tmem_ld16(trow + (uint32_t)(256 + cq * 32), v0);
      tmem_ld16(trow + (uint32_t)(256 + cq * 32 + 16), v1);
      tmem_wait_ld(v0);
      tmem_wait_ld(v1);
      float k[32];
      fold16(v0, cs + 256 + cq * 32, bq + 256 + cq * 32, rstd, nmr, k);
      fold16(v1, cs + 256 + cq * 32 + 16, bq + 256 + cq * 32 + 16, rstd, nmr, k + 16);
#pragma unroll
      for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(sX + (cq * 4 + g) * TK_PLANE + row * 16) = pack8(k + 8 * g);
    }
    __syncthreads();
    // ---- context = softmax . V -> operand planes (the x planes are dead: the qkv MMAs completed)
#pragma unroll
    for (int hd = 0; hd < 2; ++hd) {
      float ctx[16];
#pragma unroll
      for (int d = 0; d < 16; ++d) ctx[d] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint8_t* vc = sX + (cq * 4 + hd * 2) * TK_PLANE + (j * 32 + lane) * 16;
        float vv[16];
        unpack8(*reinterpret_cast<const uint4*>(vc), vv);
        unpack8(*reinterpret_cast<const uint4*>(vc + TK_PLANE), vv + 8);
#pragma unroll
        for (int d = 0; d < 16; ++d) ctx[d] = fmaf(pr[hd][j], vv[d], ctx[d]);
      }
      *reinterpret_cast<uint4*>(sA + (cq * 4 + hd * 2) * TK_PLANE + row * 16) = pack8(ctx);
      *reinterpret_cast<uint4*>(sA + (cq * 4 + hd * 2 + 1) * TK_PLANE + row * 16) = pack8(ctx + 8);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- out_proj at TMEM columns [384, 512)
    if (warp == 0) {
      tc_fence_after();
      gemm_k128(tmem + 384, a32, wo_32, 128, 0, 128, true);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      uint32_t v0[16], v1[16];
      tmem_ld16(trow + (uint32_t)(384 + cq * 32), v0);
      tmem_ld16(trow + (uint32_t)(384 + cq * 32 + 16), v1);
      tmem_wait_ld(v0);
      tmem_wait_ld(v1);
      if (live) {
        uint4* op = reinterpret_cast<uint4*>(out + orow);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float x8[8], o8[8];
          unpack8(xc[g], x8);
          const float4 b0 = *reinterpret_cast<const float4*>(bo + cq * 32 + 8 * g), b1 = *reinterpret_cast<const float4*>(bo + cq * 32 + 8 * g + 4);
          const uint32_t* vv = g < 2 ? v0 + 8 * g : v1 + 8 * (g - 2);
          o8[0] = x8[0] + (__uint_as_float(vv[0]) + b0.x); o8[1] = x8[1] + (__uint_as_float(vv[1]) + b0.y);
          o8[2] = x8[2] + (__uint_as_float(vv[2]) + b0.z); o8[3] = x8[3] + (__uint_as_float(vv[3]) + b0.w);
          o8[4] = x8[4] + (__uint_as_float(vv[4]) + b1.x); o8[5] = x8[5] + (__uint_as_float(vv[5]) + b1.y);
          o8[6] = x8[6] + (__uint_as_float(vv[6]) + b1.z); o8[7] = x8[7] + (__uint_as_float(vv[7]) + b1.w);
          op[g] = pack8(o8);
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                 // TMEM, planes and statistics are free for the next tile
    tc_fence_after();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

template <bool TANH>
__global__ void __launch_bounds__(TK_THREADS, 1) k_token_ffn(const __nv_bfloat16* __restrict__ xin, long rows,
                                                             const uint8_t* __restrict__ wblob, const float* __restrict__ pblob,
                                                             __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8);
  float* par = reinterpret_cast<float*>(smem + TF_S_PAR);
  float2* stat = reinterpret_cast<float2*>(smem + TF_S_STAT);
  uint8_t* sW = smem + TF_S_W;
  uint8_t* sA = smem + TF_S_A;
  uint8_t* sH = smem + TF_S_H;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cq = warp >> 2;
  const int row = (warp & 3) * 32 + lane;

  for (int i = tid; i < TF_WBYTES / 16; i += TK_THREADS) reinterpret_cast<uint4*>(sW)[i] = __ldg(reinterpret_cast<const uint4*>(wblob) + i);
  for (int i = tid; i < TF_PF; i += TK_THREADS) par[i] = __ldg(pblob + i);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float* cs = par;
  const float* b0 = par + 256;
  const float* b2 = par + 512;
  const uint32_t w0_32 = smem_u32(sW), w2_32 = w0_32 + TF_W0, a32 = smem_u32(sA), h32 = smem_u32(sH);
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t phase = 0;
  const long tiles = (rows + 127) / 128;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);

  uint4 nx[4];
  bool nlive = false;
  if ((long)blockIdx.x < tiles) {
    const long r = (long)blockIdx.x * 128 + row;
    nlive = r < rows;
#pragma unroll
    for (int g = 0; g < 4; ++g) nx[g] = nlive ? __ldg(reinterpret_cast<const uint4*>(xin + r * 128 + cq * 32) + g) : zero4;
  }
  for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    uint4 xc[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) xc[g] = nx[g];
    const bool live = nlive;
    const long orow = (tile * 128 + row) * 128 + cq * 32;
    put_cells_stats(xc, sA, cq, row, stat);
    if (tile + gridDim.x < tiles) {
      const long r = (tile + gridDim.x) * 128 + row;
      nlive = r < rows;
#pragma unroll
      for (int g = 0; g < 4; ++g) nx[g] = nlive ? __ldg(reinterpret_cast<const uint4*>(xin + r * 128 + cq * 32) + g) : zero4;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    // ---- hidden = x . g0^T (N = 256) at TMEM columns [128, 384)
    if (warp == 0) {
      tc_fence_after();
      gemm_k128(tmem + 128, a32, w0_32, 256, 0, 256, true);
      umma_commit(bar);
    }
    float mean, rstd;
    row_stats(stat, row, mean, rstd);
    const float nmr = -mean * rstd;
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- GELU(hidden) in two halves of 128 channels -> planes (half 0: sH, half 1: the dead x planes), ffn2 accumulates
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {
      uint32_t v0[16], v1[16];
      const int c0 = hh * 128 + cq * 32;
      tmem_ld16(trow + (uint32_t)(128 + c0), v0);
      tmem_ld16(trow + (uint32_t)(128 + c0 + 16), v1);
      tmem_wait_ld(v0);
      tmem_wait_ld(v1);
      float h[32];
      fold16(v0, cs + c0, b0 + c0, rstd, nmr, h);
      fold16(v1, cs + c0 + 16, b0 + c0 + 16, rstd, nmr, h + 16);
#pragma unroll
      for (int i = 0; i < 32; ++i) h[i] = TANH ? gelu_tanh_fast(h[i]) : gelu_erf_fast(h[i]);
      uint8_t* dst = hh == 0 ? sH : sA;
#pragma unroll
      for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(dst + (cq * 4 + g) * TK_PLANE + row * 16) = pack8(h + 8 * g);
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      if (warp == 0) {
        tc_fence_after();
        gemm_k128(tmem, hh == 0 ? h32 : a32, w2_32 + (uint32_t)(hh * 16) * 128u * 16u, 128, 0, 128, hh == 0);
        if (hh == 1) umma_commit(bar);
      }
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      uint32_t v0[16], v1[16];
      tmem_ld16(trow + (uint32_t)(cq * 32), v0);
      tmem_ld16(trow + (uint32_t)(cq * 32 + 16), v1);
      tmem_wait_ld(v0);
      tmem_wait_ld(v1);
      if (live) {
        uint4* op = reinterpret_cast<uint4*>(out + orow);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float x8[8], o8[8];
          unpack8(xc[g], x8);
          const float4 ba = *reinterpret_cast<const float4*>(b2 + cq * 32 + 8 * g), bb = *reinterpret_cast<const float4*>(b2 + cq * 32 + 8 * g + 4);
          const uint32_t* vv = g < 2 ? v0 + 8 * g : v1 + 8 * (g - 2);
          o8[0] = x8[0] + (__uint_as_float(vv[0]) + ba.x); o8[1] = x8[1] + (__uint_as_float(vv[1]) + ba.y);
          o8[2] = x8[2] + (__uint_as_float(vv[2]) + ba.z); o8[3] = x8[3] + (__uint_as_float(vv[3]) + ba.w);
          o8[4] = x8[4] + (__uint_as_float(vv[4]) + bb.x); o8[5] = x8[5] + (__uint_as_float(vv[5]) + bb.y);
          o8[6] = x8[6] + (__uint_as_float(vv[6]) + bb.z); o8[7] = x8[7] + (__uint_as_float(vv[7]) + bb.w);
          op[g] = pack8(o8);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---------------------------------------------------------------------------------------------------------------------
// k_align_tokens: the four expert feature maps (NCHW fp32 as they come out of the cache, 180 / 180 / 64 / 180 channels) ->
// aligned bf16 tokens [B][4][HW][128] = align_layers[e](feat_e)  (1x1 conv + bias, large_kernel_attention.py:344-358), in ONE
// kernel instead of four NCHW -> NHWC bf16 conversions (0.21 ms) + a grouped tcgen05 1x1 conv (0.11 ms).  A tile is 128 pixels
// of one expert: lanes run along the pixels (coalesced 128-byte reads of a channel row), a thread gathers 8 channels of its
// pixel and stores them as one 16-byte cell of the K-major operand planes, so the layout change costs no extra pass.  The
// expert's weights (48 KB, K padded to 192) stay in shared memory; the next tile's values are loaded into registers before
// the current tile's epilogue, and its MMAs (<= 12 K steps) run under the loads of the tile after.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int AL_KG = 24;                         // 8-channel groups of the padded K = 192
constexpr int AL_W = AL_KG * 128 * 16;            // 49,152 B of one expert  [kg 24][n 128][8]
constexpr int AL_S_BIAS = 64;                     // fp32 [4][128]
constexpr int AL_S_W = 4096;
constexpr int AL_S_A = AL_S_W + AL_W;
constexpr int AL_SMEM = AL_S_A + AL_KG * TK_PLANE;   // 102,400 B
constexpr int AL_THREADS = 256;                   // two CTAs per SM (2 x 100 KB of shared memory, 2 x 128 TMEM columns, 128 registers)
constexpr int AL_CPT = AL_KG / 2;                 // cells per thread: kg = slot, slot + 2, ...

struct AlignArgs {
  const float* feat[4];
  int C[4];
};

__global__ void __launch_bounds__(AL_THREADS, 2) k_align_tokens(const AlignArgs fa, int B, int HW, const uint8_t* __restrict__ wblob,
                                                                const float* __restrict__ bias, __nv_bfloat16* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8);
  float* sb = reinterpret_cast<float*>(smem + AL_S_BIAS);
  uint8_t* sW = smem + AL_S_W;
  uint8_t* sA = smem + AL_S_A;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int px = tid & 127, slot = tid >> 7;        // load role: pixel of the tile, first 8-channel group
  const int row = (warp & 3) * 32 + lane, ch = warp >> 2;   // epilogue role: TMEM lane, 64-column half

  for (int i = tid; i < 512; i += AL_THREADS) sb[i] = __ldg(bias + i);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t w32 = smem_u32(sW), a32 = smem_u32(sA);
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  const uint32_t hi = desc_hi(128), id128 = idesc_bf16_m128(128);
  const int tpi = (HW + 127) / 128;
  const long per_e = (long)B * tpi, tiles = 4 * per_e;   // expert-major tile order: a CTA changes weights at most 4 times

  // this thread's 8-channel cells of a tile, packed to bf16 (zero beyond the expert's channels and beyond the image)
  auto load_cells = [&](long tile, uint4 (&c)[AL_CPT]) {
    const int e = (int)(tile / per_e);
    const long r = tile - (long)e * per_e;
    const int b = (int)(r / tpi), p = (int)(r - (long)b * tpi) * 128 + px;
    const int Ce = fa.C[e];
    const float* f = fa.feat[e] + (long)b * Ce * HW + p;
    const bool in = p < HW;
#pragma unroll
    for (int q = 0; q < AL_CPT; ++q) {
      const int c0 = (slot + 2 * q) * 8;
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = (in && c0 + k < Ce) ? __ldg(f + (long)(c0 + k) * HW) : 0.f;
      c[q] = pack8(v);
    }
  };
  uint4 nc[AL_CPT];
  if ((long)blockIdx.x < tiles) load_cells(blockIdx.x, nc);
  uint32_t phase = 0;
  int cur_e = -1;
  for (long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int e = (int)(tile / per_e);
    if (e != cur_e) {                                // the previous tile's MMAs completed (waited for below)
      for (int i = tid; i < AL_W / 16; i += AL_THREADS) reinterpret_cast<uint4*>(sW)[i] = __ldg(reinterpret_cast<const uint4*>(wblob + (long)e * AL_W) + i);
      cur_e = e;
    }
#pragma unroll
    for (int q = 0; q < AL_CPT; ++q) *reinterpret_cast<uint4*>(sA + (slot + 2 * q) * TK_PLANE + px * 16) = nc[q];
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      const int groups = (fa.C[e] + 63) / 64;        // K steps of 16 in groups of four
      for (int g = 0; g < groups; ++g)
        umma_taps_1x4(tmem, desc_lo(a32 + (uint32_t)(g * 8) * TK_PLANE, TK_PLANE), desc_lo(w32 + (uint32_t)(g * 8) * 128u * 16u, 128u * 16u), id128,
                      g == 0 ? 0u : 1u, hi, hi, (2u * TK_PLANE) >> 4, 2u * 128u, 0u);
      umma_commit(bar);
    }
    if (tile + gridDim.x < tiles) load_cells(tile + gridDim.x, nc);   // in flight during the MMAs and the epilogue
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      const long r = tile - (long)e * per_e;
      const int b = (int)(r / tpi), p = (int)(r - (long)b * tpi) * 128 + row;
      uint4* op = reinterpret_cast<uint4*>(out + (((long)b * 4 + e) * HW + (p < HW ? p : 0)) * 128 + ch * 64);
      const float* bb = sb + e * 128 + ch * 64;
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t v0[16], v1[16];
        tmem_ld16(trow + (uint32_t)(ch * 64 + c2 * 32), v0);
        tmem_ld16(trow + (uint32_t)(ch * 64 + c2 * 32 + 16), v1);
        tmem_wait_ld(v0);
        tmem_wait_ld(v1);
        if (p < HW) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t* vv = g < 2 ? v0 + 8 * g : v1 + 8 * (g - 2);
            float o8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o8[k] = __uint_as_float(vv[k]) + bb[c2 * 32 + 8 * g + k];
            op[c2 * 4 + g] = pack8(o8);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();                                 // accumulator read, planes free
    tc_fence_after();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}
}  // namespace

extern "C" size_t ffsr_token_attn_weight_bytes(void) { return (size_t)TA_WBYTES; }
extern "C" size_t ffsr_token_attn_param_floats(void) { return (size_t)TA_PF; }
extern "C" size_t ffsr_token_ffn_weight_bytes(void) { return (size_t)TF_WBYTES; }
extern "C" size_t ffsr_token_ffn_param_floats(void) { return (size_t)TF_PF; }

// x, out: bf16 [B][4][HW][128] (out may alias x: a tile is read whole before it is written); blobs: pipeline.pack_token_attn
extern "C" int ffsr_token_attn_chain(const void* x, int B, int HW, const void* wblob, const float* pblob, void* out, cudaStream_t stream) {
  FFSR_REQUIRE(x && wblob && pblob && out && B > 0 && HW > 0, FFSR_ERR_ARG, "token_attn_chain: bad argument");
  FFSR_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)wblob % 16) == 0, FFSR_ERR_ALIGN,
               "token_attn_chain: 16-byte alignment required");
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_token_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM);
    attr = true;
  }
  const long tiles = (long)B * ((HW + 31) / 32);
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  k_token_attn<<<grid, TK_THREADS, TA_SMEM, stream>>>((const __nv_bfloat16*)x, B, HW, (const uint8_t*)wblob, pblob, (__nv_bfloat16*)out);
  return ffsr_check_launch("token_attn_chain");
}

// x, out: bf16 [rows][128]; blobs: pipeline.pack_token_ffn
extern "C" int ffsr_token_ffn_chain(const void* x, long rows, const void* wblob, const float* pblob, void* out, cudaStream_t stream) {
  FFSR_REQUIRE(x && wblob && pblob && out && rows > 0, FFSR_ERR_ARG, "token_ffn_chain: bad argument");
  FFSR_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0 && ((uintptr_t)wblob % 16) == 0, FFSR_ERR_ALIGN,
               "token_ffn_chain: 16-byte alignment required");
  static bool attr = false;
  static const bool erf_forced = getenv("FFSR_TC_GELU_ERF") != nullptr;
  if (!attr) {
    cudaFuncSetAttribute(k_token_ffn<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM);
    cudaFuncSetAttribute(k_token_ffn<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM);
    attr = true;
  }
  const long tiles = (rows + 127) / 128;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  if (erf_forced)
    k_token_ffn<false><<<grid, TK_THREADS, TF_SMEM, stream>>>((const __nv_bfloat16*)x, rows, (const uint8_t*)wblob, pblob, (__nv_bfloat16*)out);
  else
    k_token_ffn<true><<<grid, TK_THREADS, TF_SMEM, stream>>>((const __nv_bfloat16*)x, rows, (const uint8_t*)wblob, pblob, (__nv_bfloat16*)out);
  return ffsr_check_launch("token_ffn_chain");
}

extern "C" size_t ffsr_align_tokens_weight_bytes(void) { return (size_t)(4 * AL_W); }

// feat[e]: fp32 NCHW [B][C[e]][HW] of expert e (C[e] <= 192); wblob / bias: pipeline.pack_align_tokens; out: bf16 [B][4][HW][128]
extern "C" int ffsr_align_tokens(const float* const* feat, const int* C, int B, int HW, const void* wblob, const float* bias, void* out,
                                 cudaStream_t stream) {
  FFSR_REQUIRE(feat && C && wblob && bias && out && B > 0 && HW > 0, FFSR_ERR_ARG, "align_tokens: bad argument");
  AlignArgs fa;
  for (int e = 0; e < 4; ++e) {
    FFSR_REQUIRE(feat[e] && C[e] > 0 && C[e] <= AL_KG * 8, FFSR_ERR_ARG, "align_tokens: expert %d: null feature map or more than 192 channels", e);
    fa.feat[e] = feat[e];
    fa.C[e] = C[e];
  }
  FFSR_REQUIRE(((uintptr_t)out % 16) == 0 && ((uintptr_t)wblob % 16) == 0, FFSR_ERR_ALIGN, "align_tokens: 16-byte alignment required");
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_align_tokens, cudaFuncAttributeMaxDynamicSharedMemorySize, AL_SMEM);
    attr = true;
  }
  const long tiles = 4L * B * ((HW + 127) / 128);
  const int grid = (int)(tiles < 2L * sm_count() ? tiles : 2L * sm_count());
  k_align_tokens<<<grid, AL_THREADS, AL_SMEM, stream>>>(fa, B, HW, (const uint8_t*)wblob, bias, (__nv_bfloat16*)out);
  return ffsr_check_launch("align_tokens");
}
