// Tile-resident edge refiner (Phase 7b, one pyramid level): the whole EdgeRefineBlock + SpatialEdgeAttention
//   idt = proj(x); o1 = GELU(conv1(x)); o2 = GELU(conv2(o1)); o3 = conv3(o2) + idt; t8 = GELU(attn0(o3));
//   at = sigmoid(attn2(t8)); out = o3 * at                       (reference src/models/edge_enhancement.py:69-118)
// in ONE kernel: the 32-channel intermediates never leave the SM.  Replaces six k_conv_tc launches per level, each of
// which streamed its activation tensor through HBM (level 0: 0.95 ms per 2040x1356 image, every launch within 2x of its
// own HBM floor, so only keeping the intermediates on chip helps).
//
//   pass      : one CTA works on a 32 x 32-pixel compute grid (8 M-tiles of 16 rows x 8 px) at a time; conv2 / conv3 /
//               attn2 each lose one ring of the grid (their neighbours outside the grid are not computed), so a pass yields
//               the inner 26 x 26 outputs; passes overlap by 6 px (MMA work x1.5, no HBM traffic for it).
//   smem      : activations as 8-CHANNEL PLANES  plane[kg][row 0..33][col 0..33] = 16 bytes (8 bf16) with a one-pixel ring:
//               x (1 plane, TMA tile load, zero fill outside the image = the conv zero padding), two 4-plane buffers
//               A / B that the layers ping-pong between (o1 -> A, o2 -> B, o3 -> A, t8 -> B plane 0), all weights resident.
//   operands  : NO-SWIZZLE K-major UMMA descriptors straight onto the planes: the M = 128 operand of tap (dy, dx) starts
//               at ((row0 + dy) * 34 + col0 + dx) * 16 B, SBO = 34 * 16 B (next image row), LBO = plane stride (next 8
//               channels).  Verified on hardware: tools/experiments/umma_noswizzle.cu.  conv1 has 3 input channels in one
//               plane: its K = 16 step reads that plane twice (LBO = 0) against zero weights for k >= 8.
//   warps     : 0 = TMA producer, 1..3 = MMA issuers (tile j belongs to warp 1 + j % 3; issue-bound, see conv_tc.cu),
//               4..19 = epilogue (group j % 4 takes tile j; warp w reads TMEM lane quarter w % 4): tcgen05.ld -> bias /
//               GELU / residual -> zero outside the image -> st.shared into the next layer's planes.  Layers are separated
//               by fence.proxy.async + __syncthreads; accumulators (8 x 32 TMEM columns) are reused by every layer.
//   tail      : attn2 (8 -> 1, 72 MACs per pixel) and the o3 * at product run on the CUDA cores from shared memory.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_ptx.cuh"
#include "../../include/ffsr_b200.h"

namespace {
using namespace tcx;

constexpr int EC_GW = 32, EC_GH = 32;
constexpr int EC_NT = (EC_GW / 8) * (EC_GH / 16);          // 8 M-tiles per pass
constexpr int EC_PW = EC_GW + 2, EC_PH = EC_GH + 2;
constexpr int EC_SHRINK = 3;
constexpr int EC_VW = EC_GW - 2 * EC_SHRINK, EC_VH = EC_GH - 2 * EC_SHRINK;   // 26 x 26 outputs per pass
constexpr int EC_PLANE = EC_PW * EC_PH * 16;               // 18,496 B
constexpr int EC_PSTRIDE = (EC_PLANE + 127) / 128 * 128;   // 18,560 B
constexpr int EC_MMA_WARPS = 3, EC_EPI_WARPS = 16;
constexpr int EC_EPI_WARP0 = 1 + EC_MMA_WARPS;
constexpr int EC_THREADS = 32 * (EC_EPI_WARP0 + EC_EPI_WARPS);   // 640
constexpr int EC_TMEM_COLS = 256;

// bf16 weight blob, laid out as the kernel's shared memory wants it: [tap][kg][n][8 channels]
constexpr int EC_W1_OFF = 0;                               // conv1: 9 x 2 x 32 x 16 B (k >= 3 zero)
constexpr int EC_W2_OFF = EC_W1_OFF + 9 * 2 * 32 * 16;     // conv2: 9 x 4 x 32 x 16 B
constexpr int EC_W3_OFF = EC_W2_OFF + 9 * 4 * 32 * 16;     // conv3
constexpr int EC_WA_OFF = EC_W3_OFF + 9 * 4 * 32 * 16;     // attn0: 1 x 4 x 16 x 16 B (n >= 8 zero)
constexpr int EC_W_BYTES = EC_WA_OFF + 4 * 16 * 16;        // 47,104 B
// fp32 parameter blob (float offsets)
constexpr int EC_P_B1 = 0, EC_P_B2 = 32, EC_P_B3 = 64, EC_P_BA = 96, EC_P_WP = 112, EC_P_WA2 = 240, EC_P_BA2 = 312;
constexpr int EC_P_FLOATS = 320;

constexpr int EC_S_PAR = 1024;
constexpr int EC_S_X = 3072;
constexpr int EC_S_A = EC_S_X + EC_PSTRIDE;
constexpr int EC_S_B = EC_S_A + 4 * EC_PSTRIDE;
constexpr int EC_S_W = EC_S_B + 4 * EC_PSTRIDE;
constexpr int EC_SMEM = EC_S_W + EC_W_BYTES + 1024;        // + slack for the 1024-byte alignment of the base

struct EcArgs {
  int N, H, W;
  int tiles_x, tiles_y, total;
  const uint8_t* wblob;
  const float* pblob;
  const float* level_w;
  int level;
  __nv_bfloat16* dst;
  long long dst_sN, dst_sY, dst_sX;
  float* attn_out;
  int mode;        // 0: dst = o3 * at * softmax(level_w)[level];  1: dst = o3, attn_out = at
  int gelu_tanh;
};

template <bool TANH>
__device__ __forceinline__ float gelu_sel2(float x) { return TANH ? gelu_tanh_fast(x) : gelu_erf_fast(x); }

// 16 accumulator columns -> + bias (-> + projected identity) -> activation -> zero outside the image -> two 16-byte cells
template <int MODE /*0 GELU, 1 conv3 + identity*/, bool TANH>
__device__ __forceinline__ void ec_chunk(const uint32_t (&v)[16], const float* __restrict__ bias16, const float* __restrict__ wp16,
                                         float x0, float x1, float x2, bool inside, uint8_t* cell0, uint8_t* cell1) {
  float f[16];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 b4 = *reinterpret_cast<const float4*>(bias16 + 4 * k);
    f[4 * k] = __uint_as_float(v[4 * k]) + b4.x;
    f[4 * k + 1] = __uint_as_float(v[4 * k + 1]) + b4.y;
    f[4 * k + 2] = __uint_as_float(v[4 * k + 2]) + b4.z;
    f[4 * k + 3] = __uint_as_float(v[4 * k + 3]) + b4.w;
  }
  if (MODE == 0) {
#pragma unroll
    for (int k = 0; k < 16; ++k) f[k] = gelu_sel2<TANH>(f[k]);
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 w4 = *reinterpret_cast<const float4*>(wp16 + 4 * k);      // proj weights of this channel + its bias
      f[k] += fmaf(w4.x, x0, fmaf(w4.y, x1, fmaf(w4.z, x2, w4.w)));
    }
  }
  uint4 c0, c1;
  if (inside) {
    c0 = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
    c1 = make_uint4(pack_bf16(f[8], f[9]), pack_bf16(f[10], f[11]), pack_bf16(f[12], f[13]), pack_bf16(f[14], f[15]));
  } else {
    c0 = make_uint4(0u, 0u, 0u, 0u);
    c1 = c0;
  }
  *reinterpret_cast<uint4*>(cell0) = c0;
  *reinterpret_cast<uint4*>(cell1) = c1;
}

template <bool TANH>
__global__ void __launch_bounds__(EC_THREADS, 1) k_edge_chain(const __grid_constant__ CUtensorMap tmX, const EcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* xfull = reinterpret_cast<uint64_t*>(smem);
  uint64_t* tfull = xfull + 1;                                  // [EC_NT]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + EC_NT);
  float* par = reinterpret_cast<float*>(smem + EC_S_PAR);
  uint8_t* sX = smem + EC_S_X;
  uint8_t* sA = smem + EC_S_A;
  uint8_t* sB = smem + EC_S_B;
  uint8_t* sW = smem + EC_S_W;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // one-time: planes to zero (rings and never-written cells must be finite), weights and parameters resident
  for (int i = tid; i < (EC_S_W - EC_S_X) / 16; i += EC_THREADS) reinterpret_cast<uint4*>(smem + EC_S_X)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < EC_W_BYTES / 16; i += EC_THREADS) reinterpret_cast<uint4*>(sW)[i] = __ldg(reinterpret_cast<const uint4*>(a.wblob) + i);
  for (int i = tid; i < EC_P_FLOATS; i += EC_THREADS) par[i] = __ldg(a.pblob + i);
  if (tid == 0) {
    tma_prefetch_desc(&tmX);
    mbar_init(xfull, 1);
    for (int j = 0; j < EC_NT; ++j) mbar_init(&tfull[j], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(EC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  float lw = 1.f;
  if (a.mode == 0) {
    const float l0 = a.level_w[0], l1 = a.level_w[1], l2 = a.level_w[2];
    const float m = fmaxf(l0, fmaxf(l1, l2));
    const float e0 = expf(l0 - m), e1 = expf(l1 - m), e2 = expf(l2 - m);
    lw = (a.level == 0 ? e0 : (a.level == 1 ? e1 : e2)) / ((e0 + e1) + e2);
  }

  const uint32_t a_hi = desc_hi(EC_PW * 16), b_hi = desc_hi(128);
  const uint32_t sX32 = smem_u32(sX), sA32 = smem_u32(sA), sB32 = smem_u32(sB), sW32 = smem_u32(sW);

  uint32_t it = 0;
  for (int t = blockIdx.x; t < a.total; t += gridDim.x, ++it) {
    const int tx = t % a.tiles_x, ty = (t / a.tiles_x) % a.tiles_y, n = t / (a.tiles_x * a.tiles_y);
    const int gx0 = tx * EC_VW - EC_SHRINK, gy0 = ty * EC_VH - EC_SHRINK;   // image coordinates of grid pixel (0, 0)
    __syncthreads();                                   // the previous pass is done with x / A / B
    if (warp == 0) tma_load_4d_expect(sX, &tmX, xfull, (uint32_t)EC_PLANE, 0, gx0 - 1, gy0 - 1, n);

    for (int L = 0; L < 4; ++L) {
      if (warp >= 1 && warp < EC_EPI_WARP0) {
        // ------------------------------------------------ MMA issuers
        if (L == 0) mbar_wait(xfull, it & 1);
        tc_fence_after();
        const uint32_t src = L == 0 ? sX32 : (L == 2 ? sB32 : sA32);
        const uint32_t lbo_a = L == 0 ? 0u : (uint32_t)EC_PSTRIDE;
        const uint32_t wofs = L == 0 ? EC_W1_OFF : (L == 1 ? EC_W2_OFF : (L == 2 ? EC_W3_OFF : EC_WA_OFF));
        const int nout = L == 3 ? 16 : 32;
        const uint32_t kg = L == 0 ? 2u : 4u;
        const uint32_t idesc = idesc_bf16_m128(nout);
        const uint32_t a_ks = (uint32_t)(2 * EC_PSTRIDE) >> 4, b_ks = (uint32_t)(2 * nout), b_tap = kg * (uint32_t)nout;
        for (int j = warp - 1; j < EC_NT; j += EC_MMA_WARPS) {
          const int r0 = (j >> 2) * 16, c0 = (j & 3) * 8;
          const uint32_t tmem_d = tmem_base + (uint32_t)(j * 32);
          if (L < 3) {
#pragma unroll 1
            for (int dy = 0; dy < 3; ++dy) {
              const uint32_t al = desc_lo(src + (uint32_t)(((r0 + dy) * EC_PW + c0) * 16), lbo_a);
              const uint32_t bl = desc_lo(sW32 + wofs + (uint32_t)(dy * 3) * b_tap * 16u, (uint32_t)nout * 16u);
              if (L == 0) umma_taps_3x1(tmem_d, al, bl, idesc, dy > 0 ? 1u : 0u, a_hi, b_hi, a_ks, b_ks, b_tap);
              else umma_taps_3x2(tmem_d, al, bl, idesc, dy > 0 ? 1u : 0u, a_hi, b_hi, a_ks, b_ks, b_tap);
            }
          } else {
            const uint32_t al = desc_lo(src + (uint32_t)(((r0 + 1) * EC_PW + c0 + 1) * 16), lbo_a);
            const uint32_t bl = desc_lo(sW32 + wofs, (uint32_t)nout * 16u);
            umma_taps_1x2(tmem_d, al, bl, idesc, 0u, a_hi, b_hi, a_ks, b_ks, b_tap);
          }
          umma_commit(&tfull[j]);
        }
      } else if (warp >= EC_EPI_WARP0) {
        // ------------------------------------------------ epilogue
        const int wq = warp & 3, cg = (warp - EC_EPI_WARP0) >> 2;
        const uint32_t par_t = (it * 4u + (uint32_t)L) & 1u;
        uint8_t* dstp = (L == 0 || L == 2) ? sA : sB;
        for (int j = cg; j < EC_NT; j += 4) {
          const int m = wq * 32 + lane;
          const int r = (j >> 2) * 16 + (m >> 3), c = (j & 3) * 8 + (m & 7);
          const int p = (r + 1) * EC_PW + (c + 1);
          const int Y = gy0 + r, X = gx0 + c;
          const bool inside = (Y >= 0) && (Y < a.H) && (X >= 0) && (X < a.W);
          const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(j * 32);
          mbar_wait(&tfull[j], par_t);
          tc_fence_after();
          uint8_t* cell = dstp + p * 16;
          if (L < 3) {
            uint32_t v0[16], v1[16];
            tmem_ld16(taddr, v0);
            tmem_ld16(taddr + 16, v1);
            tmem_wait_ld(v0);
            tmem_wait_ld(v1);
            if (L == 2) {
              const uint4 xv = *reinterpret_cast<const uint4*>(sX + p * 16);
              const float2 x01 = unpack_bf16(xv.x), x23 = unpack_bf16(xv.y);
              ec_chunk<1, TANH>(v0, par + EC_P_B3, par + EC_P_WP, x01.x, x01.y, x23.x, inside, cell, cell + EC_PSTRIDE);
              ec_chunk<1, TANH>(v1, par + EC_P_B3 + 16, par + EC_P_WP + 64, x01.x, x01.y, x23.x, inside, cell + 2 * EC_PSTRIDE, cell + 3 * EC_PSTRIDE);
            } else {
              const float* b = par + (L == 0 ? EC_P_B1 : EC_P_B2);
              ec_chunk<0, TANH>(v0, b, nullptr, 0.f, 0.f, 0.f, inside, cell, cell + EC_PSTRIDE);
              ec_chunk<0, TANH>(v1, b + 16, nullptr, 0.f, 0.f, 0.f, inside, cell + 2 * EC_PSTRIDE, cell + 3 * EC_PSTRIDE);
            }
          } else {
            uint32_t v0[16];
            tmem_ld16(taddr, v0);
            tmem_wait_ld(v0);
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = gelu_sel2<TANH>(__uint_as_float(v0[k]) + par[EC_P_BA + k]);
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (inside) o = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
            *reinterpret_cast<uint4*>(cell) = o;
          }
        }
        fence_proxy_async();                           // st.shared above -> operand reads of the next layer's MMAs
      }
      tc_fence_before();
      __syncthreads();
    }

    // ---------------------------------------------------- tail: attn2 + product on the CUDA cores
    if (warp >= EC_EPI_WARP0) {
      for (int idx = tid - 32 * EC_EPI_WARP0; idx < EC_VW * EC_VH; idx += 32 * EC_EPI_WARPS) {
        const int vy = idx / EC_VW, vx = idx - vy * EC_VW;
        const int Y = ty * EC_VH + vy, X = tx * EC_VW + vx;
        if (Y >= a.H || X >= a.W) continue;
        const int p = (EC_SHRINK + vy + 1) * EC_PW + (EC_SHRINK + vx + 1);
        float acc = par[EC_P_BA2];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int q = p + (tap / 3 - 1) * EC_PW + (tap % 3 - 1);
          const uint4 u = *reinterpret_cast<const uint4*>(sB + q * 16);
          const float4 w0 = *reinterpret_cast<const float4*>(par + EC_P_WA2 + tap * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(par + EC_P_WA2 + tap * 8 + 4);
          const float2 t01 = unpack_bf16(u.x), t23 = unpack_bf16(u.y), t45 = unpack_bf16(u.z), t67 = unpack_bf16(u.w);
          acc = fmaf(w0.x, t01.x, acc); acc = fmaf(w0.y, t01.y, acc); acc = fmaf(w0.z, t23.x, acc); acc = fmaf(w0.w, t23.y, acc);
          acc = fmaf(w1.x, t45.x, acc); acc = fmaf(w1.y, t45.y, acc); acc = fmaf(w1.z, t67.x, acc); acc = fmaf(w1.w, t67.y, acc);
        }
        const float at = sigmoid_acc(acc);
        __nv_bfloat16* o = a.dst + (long long)n * a.dst_sN + (long long)Y * a.dst_sY + (long long)X * a.dst_sX;
        if (a.mode == 0) {
#pragma unroll
          for (int kgp = 0; kgp < 4; ++kgp) {
            const uint4 u = *reinterpret_cast<const uint4*>(sA + kgp * EC_PSTRIDE + p * 16);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
            uint32_t ow[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f2 = unpack_bf16(w[k]);
              ow[k] = pack_bf16((f2.x * at) * lw, (f2.y * at) * lw);
            }
            reinterpret_cast<uint4*>(o)[kgp] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          }
        } else {
#pragma unroll
          for (int kgp = 0; kgp < 4; ++kgp) reinterpret_cast<uint4*>(o)[kgp] = *reinterpret_cast<const uint4*>(sA + kgp * EC_PSTRIDE + p * 16);
          a.attn_out[((long long)n * a.H + Y) * a.W + X] = at;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(EC_TMEM_COLS));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn ec_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
}  // namespace

extern "C" size_t ffsr_edge_chain_weight_bytes(void) { return (size_t)EC_W_BYTES; }
extern "C" size_t ffsr_edge_chain_param_floats(void) { return (size_t)EC_P_FLOATS; }

extern "C" int ffsr_edge_refiner_chain(const void* x, int N, int H, int W, const void* wblob, const float* pblob,
                                       const float* level_w, int level, void* dst, long long dst_sN, long long dst_sY,
                                       long long dst_sX, float* attn_out, int mode, cudaStream_t stream) {
  FFSR_REQUIRE(x && wblob && pblob && dst, FFSR_ERR_ARG, "edge_refiner_chain: null pointer");
  FFSR_REQUIRE(N > 0 && H > 0 && W > 0, FFSR_ERR_ARG, "edge_refiner_chain: bad shape");
  FFSR_REQUIRE(mode == 0 ? (level_w != nullptr && level >= 0 && level < 3) : (mode == 1 && attn_out != nullptr), FFSR_ERR_ARG,
               "edge_refiner_chain: mode 0 needs level weights, mode 1 an attention output");
  FFSR_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)wblob % 16) == 0 && ((uintptr_t)dst % 16) == 0 && dst_sX % 8 == 0 &&
                   dst_sY % 8 == 0 && dst_sN % 8 == 0 && dst_sX >= 32,
               FFSR_ERR_ALIGN, "edge_refiner_chain: 16-byte aligned bf16 rows required");
  EncodeTiledFn enc = ec_encode_fn();
  FFSR_REQUIRE(enc, FFSR_ERR_DRIVER, "edge_refiner_chain: cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMap tmX;
  {
    // x: bf16 [N][H][W][8] (3 real channels); one 8-channel plane tile {8, 34, 34, 1}, zero fill outside the image
    cuuint64_t dims[4] = {8, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {16, (cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
    cuuint32_t box[4] = {8, EC_PW, EC_PH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FFSR_REQUIRE(r == CUDA_SUCCESS, FFSR_ERR_DRIVER, "edge_refiner_chain: tensor map encode failed (CUresult %d)", (int)r);
  }
  EcArgs a;
  a.N = N; a.H = H; a.W = W;
  a.tiles_x = ceil_div(W, EC_VW);
  a.tiles_y = ceil_div(H, EC_VH);
  const long long total = (long long)a.tiles_x * a.tiles_y * N;
  FFSR_REQUIRE(total < (1ll << 31), FFSR_ERR_ARG, "edge_refiner_chain: too many tiles");
  a.total = (int)total;
  a.wblob = reinterpret_cast<const uint8_t*>(wblob);
  a.pblob = pblob;
  a.level_w = level_w;
  a.level = level;
  a.dst = reinterpret_cast<__nv_bfloat16*>(dst);
  a.dst_sN = dst_sN; a.dst_sY = dst_sY; a.dst_sX = dst_sX;
  a.attn_out = attn_out;
  a.mode = mode;
  static const bool erf_forced = getenv("FFSR_TC_GELU_ERF") != nullptr;
  a.gelu_tanh = erf_forced ? 0 : 1;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(k_edge_chain<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, EC_SMEM);
    cudaFuncSetAttribute(k_edge_chain<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, EC_SMEM);
  }
  const int grid = a.total < num_sms ? a.total : num_sms;
  if (a.gelu_tanh) k_edge_chain<true><<<grid, EC_THREADS, EC_SMEM, stream>>>(tmX, a);
  else k_edge_chain<false><<<grid, EC_THREADS, EC_SMEM, stream>>>(tmX, a);
  return ffsr_check_launch("edge_refiner_chain");
}
