// Generic fp32 convolution (1x1 / 3x3, zero pad, stride 1) on CUDA cores with fused
// epilogues.  This is the fp32-precision path (<=1e-4 of the reference) and the path for
// layers too small for the tensor-core kernel; the bf16 hot layers go through conv_tc.cu.
//
// Tiling: CTA = 8x16 output pixels x CT output channels, 256 threads.  Warp w owns the
// cout group w (CT/8 channels), lane (r,q) owns pixels (r, q+4j), j=0..3, so that for a
// fixed tap the 32 lanes read 32 distinct shared-memory banks (row stride 20) and all lanes
// of a warp read the same weights (broadcast).
#include "common.cuh"
#include "../../include/ffsr_b200.h"

namespace {
constexpr int TH = 8, TW = 16, PS = 20;

__device__ __forceinline__ float ld_any(const void* base, int dtype, long long off) {
  return dtype == FFSR_DT_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[off])
                               : reinterpret_cast<const float*>(base)[off];
}

template <int KS>
struct ConvGeom {
  static constexpr int HALO = KS / 2;
  static constexpr int PH = TH + 2 * HALO;
  static constexpr int PW = TW + 2 * HALO;
  static constexpr int PLANE = PH * PS + 1;
  static constexpr int TAPS = KS * KS;
  static constexpr int CK = (KS == 3) ? 16 : 32;   // input channels per shared-memory stage
};

template <int CT, int KS, bool IN_NCHW>
__global__ void __launch_bounds__(256) k_conv_ffma(const ffsr_conv_params p) {
  using G = ConvGeom<KS>;
  constexpr int CO_T = CT / 8;
  extern __shared__ __align__(16) float smem[];
  float* sIn = smem;                         // [CK][PLANE]
  float* sW = smem + G::CK * G::PLANE;        // [TAPS][CK][CT]
  sW = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(sW) + 15) & ~uintptr_t(15));

  const int tid = threadIdx.x, lane = tid & 31, cg = tid >> 5;
  const int r = lane >> 2, q = lane & 3;
  const int tiles_x = (p.W + TW - 1) / TW;
  const int ty0 = (blockIdx.x / tiles_x) * TH, tx0 = (blockIdx.x % tiles_x) * TW;
  const int co0 = blockIdx.y * CT;
  const int n = blockIdx.z;
  const int g = n % p.groups;
  const float* __restrict__ in = reinterpret_cast<const float*>(p.in) + (long long)n * p.in_sN;
  const float* __restrict__ w = p.w + (long long)g * G::TAPS * p.Cin * p.Cout;

  float acc[4][CO_T];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < CO_T; ++k) acc[j][k] = 0.f;

  for (int c0 = 0; c0 < p.Cin; c0 += G::CK) {
    // ---- stage the input halo tile ------------------------------------------------------
    constexpr int NPIX = G::PH * G::PW;
    for (int i = tid; i < NPIX * G::CK; i += 256) {
      int ci, pix;
      if (IN_NCHW) { pix = i % NPIX; ci = i / NPIX; }
      else { ci = i % G::CK; pix = i / G::CK; }
      const int py = pix / G::PW, px = pix % G::PW;
      const int y = ty0 + py - G::HALO, x = tx0 + px - G::HALO;
      const int c = c0 + ci;
      float v = 0.f;
      if (y >= 0 && y < p.H && x >= 0 && x < p.W && c < p.Cin)
        v = in[(long long)y * p.in_sY + (long long)x * p.in_sX + (long long)c * p.in_sC];
      sIn[ci * G::PLANE + py * PS + px] = v;
    }
    // ---- stage the weights -----------------------------------------------------------------
    for (int i = tid; i < G::TAPS * G::CK * CT; i += 256) {
      const int co = i % CT, ci = (i / CT) % G::CK, tap = i / (CT * G::CK);
      const int c = c0 + ci, oc = co0 + co;
      sW[i] = (c < p.Cin && oc < p.Cout) ? w[((long long)tap * p.Cin + c) * p.Cout + oc] : 0.f;
    }
    __syncthreads();
    const int ck_n = min(G::CK, p.Cin - c0);
    for (int ci = 0; ci < ck_n; ++ci) {
      const float* sp = sIn + ci * G::PLANE + r * PS + q;
#pragma unroll
      for (int dy = 0; dy < KS; ++dy) {
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
          float a[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) a[j] = sp[dy * PS + dx + 4 * j];
          const float* wp = sW + ((dy * KS + dx) * G::CK + ci) * CT + cg * CO_T;
          float wv[CO_T];
          if (CO_T >= 4) {
#pragma unroll
            for (int k = 0; k < CO_T; k += 4) {
              const float4 t = *reinterpret_cast<const float4*>(wp + k);
              wv[k] = t.x; wv[k + 1] = t.y; wv[k + 2] = t.z; wv[k + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int k = 0; k < CO_T; ++k) wv[k] = wp[k];
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < CO_T; ++k) acc[j][k] = fmaf(a[j], wv[k], acc[j][k]);
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue ------------------------------------------------------------------------------
  const float sa = p.sa * (p.sa_ptr ? p.sa_ptr[0] : 1.0f);
  const float sb = p.sb * (p.sb_ptr ? p.sb_ptr[0] : 1.0f);
  const long long out_n = (long long)n * p.out_sN;
  const int y = ty0 + r;
  if (y >= p.H) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = tx0 + q + 4 * j;
    if (x >= p.W) continue;
#pragma unroll
    for (int k = 0; k < CO_T; ++k) {
      const int oc = co0 + cg * CO_T + k;
      if (oc >= p.Cout) continue;
      float v = acc[j][k];
      if (p.bias) v += p.bias[(long long)g * p.Cout + oc];
      if (p.epi == FFSR_EPI_LKAGATE) {
        const float xr = ld_any(p.r1, p.r1_dtype, (long long)n * p.r1_sN + (long long)y * p.r1_sY + (long long)x * p.r1_sX + oc);
        v = xr + sa * (fmaf(xr, p.ch_k[oc], p.ch_d[oc]) * sigmoid_acc(v));
      } else {
        v = apply_act(v, p.act);
        if (p.epi == FFSR_EPI_RESIDUAL) {
          v = ld_any(p.r1, p.r1_dtype, (long long)n * p.r1_sN + (long long)y * p.r1_sY + (long long)x * p.r1_sX + oc) + sa * v;
          if (p.r2) v += sb * ld_any(p.r2, p.r2_dtype, (long long)n * p.r2_sN + (long long)y * p.r2_sY + (long long)x * p.r2_sX + oc);
        }
      }
      const long long o = out_n + (long long)y * p.out_sY + (long long)x * p.out_sX + oc;
      if (p.out_dtype == FFSR_DT_BF16) reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(v);
      else reinterpret_cast<float*>(p.out)[o] = v;
    }
  }
}

template <int CT, int KS, bool IN_NCHW>
int launch_conv(const ffsr_conv_params& p, cudaStream_t stream) {
  using G = ConvGeom<KS>;
  const size_t smem = (size_t)(G::CK * G::PLANE + G::TAPS * G::CK * CT) * sizeof(float) + 16;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_conv_ffma<CT, KS, IN_NCHW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_set = true;
  }
  dim3 grid(ceil_div(p.H, TH) * ceil_div(p.W, TW), ceil_div(p.Cout, CT), p.N);
  k_conv_ffma<CT, KS, IN_NCHW><<<grid, 256, smem, stream>>>(p);
  return ffsr_check_launch("conv2d_ffma");
}

template <int KS, bool IN_NCHW>
int dispatch_ct(const ffsr_conv_params& p, cudaStream_t stream) {
  if (p.Cout > 32) return launch_conv<64, KS, IN_NCHW>(p, stream);
  if (p.Cout > 8) return launch_conv<32, KS, IN_NCHW>(p, stream);
  return launch_conv<8, KS, IN_NCHW>(p, stream);
}
}  // namespace

int ffsr_conv2d_tc(const ffsr_conv_params* p, cudaStream_t stream);   // conv_tc.cu

extern "C" int ffsr_conv2d(const ffsr_conv_params* pp, cudaStream_t stream) {
  FFSR_REQUIRE(pp, FFSR_ERR_ARG, "conv2d: null params");
  const ffsr_conv_params& p = *pp;
  FFSR_REQUIRE(p.in && p.w && p.out, FFSR_ERR_ARG, "conv2d: null pointer");
  FFSR_REQUIRE(p.N > 0 && p.H > 0 && p.W > 0 && p.Cin > 0 && p.Cout > 0, FFSR_ERR_ARG, "conv2d: bad shape");
  FFSR_REQUIRE(p.N <= 65535, FFSR_ERR_ARG, "conv2d: N > 65535");
  FFSR_REQUIRE(p.ksize == 1 || p.ksize == 3, FFSR_ERR_ARG, "conv2d: ksize must be 1 or 3 (got %d)", p.ksize);
  FFSR_REQUIRE(p.groups >= 1, FFSR_ERR_ARG, "conv2d: groups must be >= 1");
  FFSR_REQUIRE(p.epi == FFSR_EPI_PLAIN || p.r1, FFSR_ERR_ARG, "conv2d: residual epilogue needs r1");
  FFSR_REQUIRE(p.epi != FFSR_EPI_LKAGATE || (p.ch_k && p.ch_d), FFSR_ERR_ARG, "conv2d: LKA gate needs ch_k/ch_d");
  FFSR_REQUIRE(p.epi != FFSR_EPI_ACTGRAD || p.in_dtype == FFSR_DT_BF16, FFSR_ERR_ARG,
               "conv2d: FFSR_EPI_ACTGRAD is implemented on the tcgen05 (bf16) path only");
  if (p.in_dtype == FFSR_DT_BF16) return ffsr_conv2d_tc(pp, stream);       // tcgen05 implicit GEMM
  FFSR_REQUIRE(p.w_dtype == FFSR_DT_F32, FFSR_ERR_ARG, "conv2d: fp32 input needs fp32-packed weights");
  const bool nchw = (p.in_sX == 1 && p.in_sC != 1);
  if (p.ksize == 3) return nchw ? dispatch_ct<3, true>(p, stream) : dispatch_ct<3, false>(p, stream);
  return nchw ? dispatch_ct<1, true>(p, stream) : dispatch_ct<1, false>(p, stream);
}
