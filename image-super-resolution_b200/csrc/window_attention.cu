// (Shifted-)window multi-head attention of the DRCT-L expert -- FIRST, CUDA-core version.
//
// STATUS: parity-green on a B200 for every (dim, heads) pair of DRCT-L, shifted and not (tests/test_gpu_drct.py: fp32
// <= 2e-5, bf16 <= 2e-2 against the torch restatement of the reference block) and inside isr_b200.drct.DRCT.forward;
// not yet timed -- a tcgen05 version of the two products is the next step.
//
// Replaces, for one SwinTransformerBlock, everything between the qkv Linear and the proj Linear
// (src/models/drct/drct_arch.py:175-206 and 385-412): cyclic shift, window partition, q k^T / sqrt(dh) + relative
// position bias (+ the -100 shift mask), softmax, attn @ v, window merge, reverse shift -- as index arithmetic inside
// one kernel instead of five tensor permutations.
//
//   qkv : [B][H][W][3*C] channels-last, channel = (which in {q,k,v}) * C + head * dh + d   (the reshape of :177)
//   out : [B][H][W][C]   channel = head * dh + d, at the ORIGINAL (un-shifted) pixel
//   bias_table : [(2*ws-1)^2][heads] fp32 (relative_position_bias_table)
//
// One CTA per (window, head): K and V of the window (N = ws*ws <= 256 tokens x dh <= 128) are staged in shared memory,
// each warp then takes queries round-robin: lanes own keys for q.k^T (q broadcast from shared memory), the softmax is
// a warp reduction, and lanes own output channels for p.V (p broadcast by shuffle).
#include "common.cuh"
#include "../../include/ffsr_b200.h"

int ffsr_window_attention_tc_try(const void* qkv, long qkv_pitch, int B, int H, int W, int C, int heads, int window, int shift,
                                 const float* bias_table, void* out, long out_pitch, cudaStream_t stream);

namespace {

constexpr int WA_THREADS = 256;
constexpr int WA_MAXN = 256;        // tokens per window
constexpr int WA_MAXD = 128;        // head dim

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// region index of the shift mask (calculate_mask, drct_arch.py:353-374) along one axis of the SHIFTED frame
__device__ __forceinline__ int mask_region(int p, int n, int ws, int shift) { return p < n - ws ? 0 : (p < n - shift ? 1 : 2); }

// T: tensor element type; TS: shared-memory storage type of K / V (float, or bf16 when T is bf16 and fp32 would not fit).
// VGLOBAL: only K is staged and V is read from global memory / L2 in the p.V loop -- the fp32 fallback for the one shape
// whose fp32 K + V exceed shared memory (DRCT-L swin3: 256 tokens x head dim 122 = 2 x 126 KB).
// PITCHED: token rows of qkv / out have pitches qp / op (elements) instead of 3*C / C -- bf16 buffers of the DRCT channel
// counts (all 4 mod 8) are padded to a multiple of 8 channels for the 16-byte TMA stride rule of the tcgen05 Linears.
// The two trailing parameters are ignored otherwise (kept last so the un-pitched instantiations compile to the same code).
template <typename T, typename TS, bool VGLOBAL = false, bool PITCHED = false>
__global__ void __launch_bounds__(WA_THREADS) k_window_attn(const T* __restrict__ qkv, int B, int H, int W, int C, int heads, int ws,
                                                            int shift, const float* __restrict__ table, T* __restrict__ out,
                                                            long qp_arg, long op_arg) {
#define WA_QP (PITCHED ? (size_t)qp_arg : (size_t)(3 * C))
#define WA_OP (PITCHED ? (size_t)op_arg : (size_t)C)
  extern __shared__ __align__(16) unsigned char wa_smem[];
  const int N = ws * ws, dh = C / heads;
  const int P = dh | 1;                                        // odd pitch: lanes reading one column hit distinct banks
  TS* Ks = reinterpret_cast<TS*>(wa_smem);
  TS* Vs = Ks + (size_t)N * P;                                 // (unused when VGLOBAL)
  float* qs = reinterpret_cast<float*>(wa_smem + (((size_t)(VGLOBAL ? 1 : 2) * N * P * sizeof(TS) + 15) & ~(size_t)15));   // [8][WA_MAXD]
  int* pix = reinterpret_cast<int*>(qs + 8 * WA_MAXD);         // [N] original pixel index y*W + x of token t
  unsigned char* rid = reinterpret_cast<unsigned char*>(pix + N);   // [N] mask region of token t

  const int nwx = W / ws, nwy = H / ws;
  int wid = blockIdx.x;
  const int wx = wid % nwx; wid /= nwx;
  const int wy = wid % nwy;
  const int b = wid / nwy;
  const int h = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int t = tid; t < N; t += WA_THREADS) {
    const int ys = wy * ws + t / ws, xs = wx * ws + t % ws;    // position in the shifted frame
    int yo = ys + shift, xo = xs + shift;                      // torch.roll(x, -shift)[i] = x[(i + shift) % n]
    if (yo >= H) yo -= H;
    if (xo >= W) xo -= W;
    pix[t] = yo * W + xo;
    rid[t] = shift ? (unsigned char)(mask_region(ys, H, ws, shift) * 3 + mask_region(xs, W, ws, shift)) : 0;
  }
  __syncthreads();
  const size_t img = (size_t)b * H * W;
  for (int i = tid; i < N * dh; i += WA_THREADS) {
    const int t = i / dh, d = i - t * dh;
    const T* src = qkv + (img + pix[t]) * WA_QP + h * dh + d;
    Ks[t * P + d] = from_f<TS>(to_f<T>(src[C]));
    if (!VGLOBAL) Vs[t * P + d] = from_f<TS>(to_f<T>(src[2 * C]));
  }
  __syncthreads();

  const float scale = rsqrtf((float)dh);
  float* q = qs + warp * WA_MAXD;
  const int nk = (N + 31) >> 5;                                // key slots per lane (<= 8)
  for (int tq = warp; tq < N; tq += WA_THREADS / 32) {
    const T* qsrc = qkv + (img + pix[tq]) * WA_QP + h * dh;
    __syncwarp();
    for (int d = lane; d < dh; d += 32) q[d] = to_f<T>(qsrc[d]) * scale;
    __syncwarp();
    const int qy = tq / ws, qx = tq - qy * ws;
    const int qr = rid[tq];
    float s[8];
    float mx = -INFINITY;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      s[kk] = -INFINITY;
      const int j = lane + 32 * kk;
      if (kk < nk && j < N) {
        const TS* kr = Ks + j * P;
        float a = 0.f;
        for (int d = 0; d < dh; ++d) a = fmaf(q[d], to_f<TS>(kr[d]), a);
        const int jy = j / ws, jx = j - jy * ws;
        a += __ldg(table + ((qy - jy + ws - 1) * (2 * ws - 1) + (qx - jx + ws - 1)) * heads + h);
        if (rid[j] != qr) a += -100.0f;
        s[kk] = a;
        mx = fmaxf(mx, a);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float den = 0.f;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      s[kk] = (s[kk] == -INFINITY) ? 0.f : expf(s[kk] - mx);
      den += s[kk];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
    const float inv = 1.0f / den;

    float acc[4] = {0.f, 0.f, 0.f, 0.f};                        // output channels lane, lane+32, lane+64, lane+96
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      if (kk < nk) {
        for (int l = 0; l < 32; ++l) {
          const float pj = __shfl_sync(0xffffffffu, s[kk], l);
          const int j = l + 32 * kk;
          if (j < N) {
            if (VGLOBAL) {
              const T* vg = qkv + (img + pix[j]) * WA_QP + 2 * C + h * dh;
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                const int d = lane + 32 * m;
                if (d < dh) acc[m] = fmaf(pj, to_f<T>(vg[d]), acc[m]);
              }
            } else {
              const TS* vr = Vs + j * P;
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                const int d = lane + 32 * m;
                if (d < dh) acc[m] = fmaf(pj, to_f<TS>(vr[d]), acc[m]);
              }
            }
          }
        }
      }
    }
    T* dst = out + (img + pix[tq]) * WA_OP + h * dh;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int d = lane + 32 * m;
      if (d < dh) dst[d] = from_f<T>(acc[m] * inv);
    }
  }
}

#undef WA_QP
#undef WA_OP

size_t wa_smem_bytes(int N, int dh, size_t esz, int copies = 2) {
  const size_t P = (size_t)(dh | 1);
  return (((size_t)copies * N * P * esz + 15) & ~(size_t)15) + 8 * WA_MAXD * sizeof(float) + (size_t)N * sizeof(int) + (size_t)N + 16;
}

}  // namespace

extern "C" int ffsr_window_attention(const void* qkv, int B, int H, int W, int C, int heads, int window, int shift,
                                     const float* bias_table, void* out, int dtype, cudaStream_t stream) {
  FFSR_REQUIRE(qkv && out && bias_table, FFSR_ERR_ARG, "window_attention: null pointer");
  FFSR_REQUIRE(B > 0 && H > 0 && W > 0 && heads > 0 && C > 0 && C % heads == 0, FFSR_ERR_ARG,
               "window_attention: B=%d H=%d W=%d C=%d heads=%d", B, H, W, C, heads);
  FFSR_REQUIRE(window > 0 && window * window <= WA_MAXN && H % window == 0 && W % window == 0, FFSR_ERR_ARG,
               "window_attention: window %d needs window^2 <= %d and H, W (%d, %d) multiples of it (callers pad)", window, WA_MAXN, H, W);
  FFSR_REQUIRE(shift >= 0 && shift < window, FFSR_ERR_ARG, "window_attention: shift %d outside [0, %d)", shift, window);
  FFSR_REQUIRE(C / heads <= WA_MAXD, FFSR_ERR_ARG, "window_attention: head dim %d > %d", C / heads, WA_MAXD);
  FFSR_REQUIRE(dtype == FFSR_DT_F32 || dtype == FFSR_DT_BF16, FFSR_ERR_ARG, "window_attention: dtype %d", dtype);
  const int N = window * window, dh = C / heads;
  const size_t limit = 227 * 1024;
  const dim3 grid((unsigned)((long)B * (H / window) * (W / window)), (unsigned)heads);
  const size_t s32 = wa_smem_bytes(N, dh, 4), s16 = wa_smem_bytes(N, dh, 2);
  if (dtype == FFSR_DT_BF16) {
    if (s32 <= limit) {
      auto k = k_window_attn<__nv_bfloat16, float>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s32);
      k<<<grid, WA_THREADS, s32, stream>>>((const __nv_bfloat16*)qkv, B, H, W, C, heads, window, shift, bias_table, (__nv_bfloat16*)out, 0L, 0L);
    } else {
      FFSR_REQUIRE(s16 <= limit, FFSR_ERR_ARG, "window_attention: K/V of a %d-token window x head dim %d do not fit in shared memory", N, dh);
      auto k = k_window_attn<__nv_bfloat16, __nv_bfloat16>;      // K / V are bf16 values already: storing them as bf16 is exact
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s16);
      k<<<grid, WA_THREADS, s16, stream>>>((const __nv_bfloat16*)qkv, B, H, W, C, heads, window, shift, bias_table, (__nv_bfloat16*)out, 0L, 0L);
    }
  } else {
    if (s32 <= limit) {
      auto k = k_window_attn<float, float>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s32);
      k<<<grid, WA_THREADS, s32, stream>>>((const float*)qkv, B, H, W, C, heads, window, shift, bias_table, (float*)out, 0L, 0L);
    } else {
      // fp32 K + V do not fit (DRCT-L swin3, head dim 122 at window 16): stage K only, read V through L2.
      // NOT yet run on hardware (written after the round's GPU budget was spent); same arithmetic as the staged path.
      const size_t s1 = wa_smem_bytes(N, dh, 4, 1);
      FFSR_REQUIRE(s1 <= limit, FFSR_ERR_ARG, "window_attention: fp32 K of a %d-token window x head dim %d needs %zu B of shared memory (> %zu)", N, dh, s1, limit);
      auto k = k_window_attn<float, float, true>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s1);
      k<<<grid, WA_THREADS, s1, stream>>>((const float*)qkv, B, H, W, C, heads, window, shift, bias_table, (float*)out, 0L, 0L);
    }
  }
  return ffsr_check_launch("k_window_attn");
}

// Same operation on PADDED rows: qkv rows of pitch qkv_pitch >= 3*C, out rows of pitch out_pitch >= C (elements).  bf16 only
// (the layout exists for the tcgen05 Linears around it; caller: isr_b200.drct with precision "bf16").
extern "C" int ffsr_window_attention_pitched(const void* qkv, long qkv_pitch, int B, int H, int W, int C, int heads, int window, int shift,
                                             const float* bias_table, void* out, long out_pitch, cudaStream_t stream) {
  FFSR_REQUIRE(qkv && out && bias_table, FFSR_ERR_ARG, "window_attention_pitched: null pointer");
  FFSR_REQUIRE(B > 0 && H > 0 && W > 0 && heads > 0 && C > 0 && C % heads == 0 && qkv_pitch >= 3L * C && out_pitch >= C, FFSR_ERR_ARG,
               "window_attention_pitched: B=%d H=%d W=%d C=%d heads=%d pitches %ld / %ld", B, H, W, C, heads, qkv_pitch, out_pitch);
  FFSR_REQUIRE(window > 0 && window * window <= WA_MAXN && H % window == 0 && W % window == 0 && shift >= 0 && shift < window &&
                   C / heads <= WA_MAXD, FFSR_ERR_ARG, "window_attention_pitched: window %d shift %d head dim %d", window, shift, C / heads);
  {
    // 16 x 16 windows: both products on tcgen05 (window_attention_tc.cu); other shapes fall through to the CUDA-core kernel
    const int r = ffsr_window_attention_tc_try(qkv, qkv_pitch, B, H, W, C, heads, window, shift, bias_table, out, out_pitch, stream);
    if (r != 0) return r == 1 ? 0 : r;
  }
  const int N = window * window, dh = C / heads;
  const size_t limit = 227 * 1024;
  const dim3 grid((unsigned)((long)B * (H / window) * (W / window)), (unsigned)heads);
  const size_t s32 = wa_smem_bytes(N, dh, 4), s16 = wa_smem_bytes(N, dh, 2);
  if (s32 <= limit) {
    auto k = k_window_attn<__nv_bfloat16, float, false, true>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s32);
    k<<<grid, WA_THREADS, s32, stream>>>((const __nv_bfloat16*)qkv, B, H, W, C, heads, window, shift, bias_table, (__nv_bfloat16*)out, qkv_pitch, out_pitch);
  } else {
    FFSR_REQUIRE(s16 <= limit, FFSR_ERR_ARG, "window_attention_pitched: K/V of a %d-token window x head dim %d do not fit in shared memory", N, dh);
    auto k = k_window_attn<__nv_bfloat16, __nv_bfloat16, false, true>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s16);
    k<<<grid, WA_THREADS, s16, stream>>>((const __nv_bfloat16*)qkv, B, H, W, C, heads, window, shift, bias_table, (__nv_bfloat16*)out, qkv_pitch, out_pitch);
  }
  return ffsr_check_launch("k_window_attn(pitched)");
}
