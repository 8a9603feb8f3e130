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

// ------------------------------------------------------------------------------------------------
// 1x1 convolution on dense channels-last fp32 rows as a register-tiled SGEMM:
//   C[M = N*H*W pixels][Cout] = A[M][Cin] * W[Cin][Cout]  (+ the same fused epilogues)
// CTA tile 128 pixels x 64 couts, K staged 16 at a time, 256 threads x (8 x 4) accumulators, A stored
// transposed in shared memory so that a thread's 8 pixels are two 16-byte reads; the next K stage is
// prefetched into registers while the current one is multiplied.
// ------------------------------------------------------------------------------------------------
constexpr int GM = 128, GN = 64, GK = 16;

__global__ void __launch_bounds__(256) k_conv1x1_gemm(const ffsr_conv_params p, long M) {
  __shared__ __align__(16) float As[GK][GM + 4];
  __shared__ __align__(16) float Bs[GK][GN];
  const int tid = threadIdx.x;
  const int tm = tid >> 4, tn = tid & 15;
  const long m0 = (long)blockIdx.x * GM;
  const int n0 = blockIdx.y * GN;
  const float* __restrict__ A = reinterpret_cast<const float*>(p.in);
  const float* __restrict__ Wt = p.w;
  const int K = p.Cin, Nc = p.Cout;
  // global->register staging: A: 128 rows x 16 k = 512 float4 (2 per thread); B: 16 k x 64 n = 256 float4 (1 per thread)
  const int a_row0 = tid >> 2, a_k4 = (tid & 3) << 2;            // second A load: row + 64
  const int b_k = tid >> 4, b_n4 = (tid & 15) << 2;
  float4 ra[2], rb;
  auto load_stage = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const long row = m0 + a_row0 + 64 * r;
      const int k = k0 + a_k4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < M) {
        const float* src = A + row * p.in_sX + k;
        if (k + 3 < K) v = *reinterpret_cast<const float4*>(src);
        else {
          if (k < K) v.x = src[0];
          if (k + 1 < K) v.y = src[1];
          if (k + 2 < K) v.z = src[2];
        }
      }
      ra[r] = v;
    }
    {
      const int k = k0 + b_k, n = n0 + b_n4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K) {
        const float* src = Wt + (long)k * Nc + n;
        if (n + 3 < Nc && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) v = *reinterpret_cast<const float4*>(src);
        else {
          if (n < Nc) v.x = src[0];
          if (n + 1 < Nc) v.y = src[1];
          if (n + 2 < Nc) v.z = src[2];
          if (n + 3 < Nc) v.w = src[3];
        }
      }
      rb = v;
    }
  };
  auto store_stage = [&]() {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = a_row0 + 64 * r;
      As[a_k4][row] = ra[r].x; As[a_k4 + 1][row] = ra[r].y; As[a_k4 + 2][row] = ra[r].z; As[a_k4 + 3][row] = ra[r].w;
    }
    *reinterpret_cast<float4*>(&Bs[b_k][b_n4]) = rb;
  };
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  load_stage(0);
  for (int k0 = 0; k0 < K; k0 += GK) {
    store_stage();
    __syncthreads();
    if (k0 + GK < K) load_stage(k0 + GK);
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][tm * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tn * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue (same semantics as k_conv_ffma) ---------------------------------------------------
  const float sa = p.sa * (p.sa_ptr ? p.sa_ptr[0] : 1.0f);
  const float sb = p.sb * (p.sb_ptr ? p.sb_ptr[0] : 1.0f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long row = m0 + tm * 8 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int oc = n0 + tn * 4 + j;
      if (oc >= Nc) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[oc];
      if (p.epi == FFSR_EPI_LKAGATE) {
        const float xr = ld_any(p.r1, p.r1_dtype, row * p.r1_sX + oc);
        v = xr + sa * (fmaf(xr, p.ch_k[oc], p.ch_d[oc]) * sigmoid_acc(v));
      } else {
        v = apply_act(v, p.act);
        if (p.epi == FFSR_EPI_RESIDUAL) {
          v = ld_any(p.r1, p.r1_dtype, row * p.r1_sX + oc) + sa * v;
          if (p.r2) v += sb * ld_any(p.r2, p.r2_dtype, row * p.r2_sX + oc);
        }
      }
      const long long o = row * p.out_sX + oc;
      if (p.out_dtype == FFSR_DT_BF16) reinterpret_cast<__nv_bfloat16*>(p.out)[o] = __float2bfloat16_rn(v);
      else reinterpret_cast<float*>(p.out)[o] = v;
    }
  }
}

// rows are dense when every image / image row continues the previous one at the pixel pitch
static bool dense_rows(long long sN, long long sY, long long sX, int H, int W) {
  return sY == (long long)W * sX && sN == (long long)H * W * sX;
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
  if (p.ksize == 1 && p.groups == 1 && p.in_sC == 1 && p.epi != FFSR_EPI_ACTGRAD && p.Cin >= 16 &&
      dense_rows(p.in_sN, p.in_sY, p.in_sX, p.H, p.W) && dense_rows(p.out_sN, p.out_sY, p.out_sX, p.H, p.W) &&
      (!p.r1 || dense_rows(p.r1_sN, p.r1_sY, p.r1_sX, p.H, p.W)) && (!p.r2 || dense_rows(p.r2_sN, p.r2_sY, p.r2_sX, p.H, p.W)) &&
      p.in_sX % 4 == 0 && ((uintptr_t)p.in % 16) == 0) {
    const long M = (long)p.N * p.H * p.W;
    dim3 grid((unsigned)((M + GM - 1) / GM), (unsigned)((p.Cout + GN - 1) / GN));
    k_conv1x1_gemm<<<grid, 256, 0, stream>>>(p, M);
    return ffsr_check_launch("conv1x1_gemm");
  }
  if (p.ksize == 3) return nchw ? dispatch_ct<3, true>(p, stream) : dispatch_ct<3, false>(p, stream);
  return nchw ? dispatch_ct<1, true>(p, stream) : dispatch_ct<1, false>(p, stream);
}
