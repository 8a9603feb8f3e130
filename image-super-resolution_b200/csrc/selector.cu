// Phase 6 selector nets as ONE kernel (fp32, CUDA cores): DynamicExpertSelector.forward, src/models/enhanced_fusion_v2.py:450-466
//   difficulty = sigmoid(conv3x3(relu(conv3x3(relu(conv3x3(r, 3->32)), 32->32)), 32->1))
//   raw        = conv1x1(relu(conv3x3(relu(conv3x3(r, 3->32)), 32->32)), 32->4)
//   gates      = sigmoid(T (raw - (0.7 - 0.5 difficulty))) / max(sum + 1e-8, 0.3)
// Seven launches before (six generic fp32 convs at ~30 % of the FMA peak on a 0.17-MPix image + the gate normalisation, 0.60 ms
// per C3 image); here a CTA keeps a 16x16-pixel tile with its halos in shared memory through both branches:
//   r 22x22x3 -> d0 20x20x32 -> d2 18x18x32 -> d4 16x16;   r -> g0 18x18x32 -> g2 16x16x32 -> g4 -> gates
// Every conv zero-pads ITS OWN input, so intermediate pixels outside the image are stored as 0.
// The 32->32 layers (92 % of the FMAs) run as register tiles of PXT pixels x 8 output channels per thread on vector loads
// (activations [pixel][36] floats: 4 input channels per load, the 36-float pitch keeps the 8 lanes of a load phase on distinct
// banks; weights [tap][ci][co]: 8 output channels per two loads, broadcast over the pixel lanes): ~9 FMAs per shared load.
// fp32 throughout: the outputs feed the expert-selection indices, which must match the reference bit for bit.
#include "common.cuh"
#include "../../include/ffsr_b200.h"

namespace {
constexpr int SL_T = 16, SL_R = SL_T + 6, SL_A = SL_T + 4, SL_B = SL_T + 2;
constexpr int SL_PS = 36;                         // floats per pixel of the 32-channel tiles
constexpr int SL_THREADS = 512;
// weight blob (floats): see isr_b200.pipeline.pack_selector
constexpr int SL_W0 = 9 * 3 * 32, SL_W2 = 9 * 32 * 32, SL_W4D = 9 * 32, SL_W4G = 32 * 4;
constexpr int SL_O_D0W = 0, SL_O_D0B = SL_O_D0W + SL_W0, SL_O_D2W = SL_O_D0B + 32, SL_O_D2B = SL_O_D2W + SL_W2,
              SL_O_D4W = SL_O_D2B + 32, SL_O_D4B = SL_O_D4W + SL_W4D, SL_O_G0W = SL_O_D4B + 4, SL_O_G0B = SL_O_G0W + SL_W0,
              SL_O_G2W = SL_O_G0B + 32, SL_O_G2B = SL_O_G2W + SL_W2, SL_O_G4W = SL_O_G2B + 32, SL_O_G4B = SL_O_G4W + SL_W4G,
              SL_BLOB = SL_O_G4B + 4;

struct SelSmem {
  float r[3][SL_R * SL_R];
  float w2[SL_W2];                                // d2 weights, then g2 weights
  float w0[2][SL_W0 + 32];                        // d0 | g0 weights + bias
  float b2[2][32];
  float w4d[SL_W4D + 4];
  float w4g[SL_W4G + 4];
  float diff[SL_T * SL_T];
  float t1[SL_A * SL_A * SL_PS];                  // d0 (20x20), then g0 (18x18)
  float t2[SL_B * SL_B * SL_PS];                  // d2 (18x18), then g2 (16x16)
};

// relu(conv3x3(r, 3 -> 32)) on an OW x OW region whose pixel (0,0) is routing-tile pixel (off, off); zero outside the image
template <int OW>
__device__ __forceinline__ void sel_conv_in(const SelSmem& s, const float* __restrict__ w0, float* __restrict__ dst, int off, int gy0,
                                            int gx0, int H, int W, int tid) {
  for (int it = tid; it < OW * OW * 4; it += SL_THREADS) {
    const int cg = it & 3, px = it >> 2;
    const int oy = px / OW, ox = px - oy * OW;
    float acc[8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(w0 + SL_W0 + cg * 8), b1 = *reinterpret_cast<const float4*>(w0 + SL_W0 + cg * 8 + 4);
      acc[0] = b0.x; acc[1] = b0.y; acc[2] = b0.z; acc[3] = b0.w; acc[4] = b1.x; acc[5] = b1.y; acc[6] = b1.z; acc[7] = b1.w;
    }
#pragma unroll 1
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float a = s.r[c][(oy + off + dy - 1) * SL_R + ox + off + dx - 1];
          const float* wp = w0 + ((dy * 3 + dx) * 3 + c) * 32 + cg * 8;
          const float4 u = *reinterpret_cast<const float4*>(wp), v = *reinterpret_cast<const float4*>(wp + 4);
          acc[0] = fmaf(a, u.x, acc[0]); acc[1] = fmaf(a, u.y, acc[1]); acc[2] = fmaf(a, u.z, acc[2]); acc[3] = fmaf(a, u.w, acc[3]);
          acc[4] = fmaf(a, v.x, acc[4]); acc[5] = fmaf(a, v.y, acc[5]); acc[6] = fmaf(a, v.z, acc[6]); acc[7] = fmaf(a, v.w, acc[7]);
        }
    const int gy = gy0 + oy, gx = gx0 + ox;
    const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
    float* d = dst + px * SL_PS + cg * 8;
    *reinterpret_cast<float4*>(d) = in ? make_float4(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f), fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f)) : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(d + 4) = in ? make_float4(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f), fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// relu(conv3x3(src, 32 -> 32) + b) : src [IW*IW][36] -> dst [OW*OW][36], OW = IW - 2; PXT pixels x 8 channels per thread
template <int IW, int PXT>
__device__ __forceinline__ void sel_conv32(const float* __restrict__ src, const float* __restrict__ w, const float* __restrict__ bias,
                                           float* __restrict__ dst, int gy0, int gx0, int H, int W, int tid) {
  constexpr int OW = IW - 2, GPR = OW / PXT;       // pixel groups per row
  static_assert(OW % PXT == 0, "row must split into whole pixel groups");
  for (int it = tid; it < OW * GPR * 4; it += SL_THREADS) {
    const int cg = it & 3, pg = it >> 2;
    const int oy = pg / GPR, ox0 = (pg - oy * GPR) * PXT;
    float acc[PXT][8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(bias + cg * 8), b1 = *reinterpret_cast<const float4*>(bias + cg * 8 + 4);
#pragma unroll
      for (int p = 0; p < PXT; ++p) {
        acc[p][0] = b0.x; acc[p][1] = b0.y; acc[p][2] = b0.z; acc[p][3] = b0.w;
        acc[p][4] = b1.x; acc[p][5] = b1.y; acc[p][6] = b1.z; acc[p][7] = b1.w;
      }
    }
#pragma unroll 1
    for (int dy = 0; dy < 3; ++dy) {
      const float* row = src + ((oy + dy) * IW + ox0) * SL_PS;
#pragma unroll 1
      for (int c4 = 0; c4 < 8; ++c4) {
        float4 a[PXT + 2];
#pragma unroll
        for (int i = 0; i < PXT + 2; ++i) a[i] = *reinterpret_cast<const float4*>(row + i * SL_PS + c4 * 4);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float* wp = w + ((dy * 3 + dx) * 32 + c4 * 4) * 32 + cg * 8;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 u = *reinterpret_cast<const float4*>(wp + k * 32), v = *reinterpret_cast<const float4*>(wp + k * 32 + 4);
#pragma unroll
            for (int p = 0; p < PXT; ++p) {
              const float av = k == 0 ? a[p + dx].x : k == 1 ? a[p + dx].y : k == 2 ? a[p + dx].z : a[p + dx].w;
              acc[p][0] = fmaf(av, u.x, acc[p][0]); acc[p][1] = fmaf(av, u.y, acc[p][1]);
              acc[p][2] = fmaf(av, u.z, acc[p][2]); acc[p][3] = fmaf(av, u.w, acc[p][3]);
              acc[p][4] = fmaf(av, v.x, acc[p][4]); acc[p][5] = fmaf(av, v.y, acc[p][5]);
              acc[p][6] = fmaf(av, v.z, acc[p][6]); acc[p][7] = fmaf(av, v.w, acc[p][7]);
            }
          }
        }
      }
    }
    const int gy = gy0 + oy;
#pragma unroll
    for (int p = 0; p < PXT; ++p) {
      const int gx = gx0 + ox0 + p;
      const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W;
      float* d = dst + (oy * OW + ox0 + p) * SL_PS + cg * 8;
      *reinterpret_cast<float4*>(d) = in ? make_float4(fmaxf(acc[p][0], 0.f), fmaxf(acc[p][1], 0.f), fmaxf(acc[p][2], 0.f), fmaxf(acc[p][3], 0.f)) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(d + 4) = in ? make_float4(fmaxf(acc[p][4], 0.f), fmaxf(acc[p][5], 0.f), fmaxf(acc[p][6], 0.f), fmaxf(acc[p][7], 0.f)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

__global__ void __launch_bounds__(SL_THREADS, 1) k_selector(const float* __restrict__ routing, int B, int H, int W,
                                                            const float* __restrict__ blob, const float* __restrict__ temperature,
                                                            float* __restrict__ diff, float* __restrict__ graw, float* __restrict__ gates) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SelSmem& s = *reinterpret_cast<SelSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int tiles_x = (W + SL_T - 1) / SL_T, tiles_y = (H + SL_T - 1) / SL_T;
  const long HW = (long)H * W;
  // weights that stay for the whole CTA
  for (int i = tid; i < SL_W0 + 32; i += SL_THREADS) {
    s.w0[0][i] = blob[SL_O_D0W + i];
    s.w0[1][i] = blob[SL_O_G0W + i];
  }
  if (tid < 32) { s.b2[0][tid] = blob[SL_O_D2B + tid]; s.b2[1][tid] = blob[SL_O_G2B + tid]; }
  for (int i = tid; i < SL_W4D + 4; i += SL_THREADS) s.w4d[i] = blob[SL_O_D4W + i];
  for (int i = tid; i < SL_W4G + 4; i += SL_THREADS) s.w4g[i] = blob[SL_O_G4W + i];
  const float T = temperature[0];

  for (int tile = blockIdx.x; tile < B * tiles_y * tiles_x; tile += gridDim.x) {
    const int b = tile / (tiles_y * tiles_x), tr = tile - b * tiles_y * tiles_x;
    const int y0 = (tr / tiles_x) * SL_T, x0 = (tr % tiles_x) * SL_T;
    __syncthreads();                                // previous tile's readers of r / t2 / diff are done
    for (int i = tid; i < 3 * SL_R * SL_R; i += SL_THREADS) {
      const int c = i / (SL_R * SL_R), p = i - c * SL_R * SL_R;
      const int gy = y0 - 3 + p / SL_R, gx = x0 - 3 + p % SL_R;
      s.r[c][p] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(routing + ((long)b * 3 + c) * HW + (long)gy * W + gx) : 0.f;
    }
    for (int i = tid; i < SL_W2 / 4; i += SL_THREADS) reinterpret_cast<float4*>(s.w2)[i] = __ldg(reinterpret_cast<const float4*>(blob + SL_O_D2W) + i);
    __syncthreads();
    // ---- difficulty branch
    sel_conv_in<SL_A>(s, s.w0[0], s.t1, 1, y0 - 2, x0 - 2, H, W, tid);
    __syncthreads();
    sel_conv32<SL_A, 3>(s.t1, s.w2, s.b2[0], s.t2, y0 - 1, x0 - 1, H, W, tid);
    __syncthreads();
    // d4 (32 -> 1, 3x3) + sigmoid on the 16x16 tile; g2's weights replace d2's meanwhile; g0 overwrites t1 (d2 is done with it)
    for (int i = tid; i < SL_W2 / 4; i += SL_THREADS) reinterpret_cast<float4*>(s.w2)[i] = __ldg(reinterpret_cast<const float4*>(blob + SL_O_G2W) + i);
    if (tid < SL_T * SL_T) {
      const int oy = tid / SL_T, ox = tid % SL_T;
      float acc = s.w4d[SL_W4D];
#pragma unroll 1
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float* ap = s.t2 + ((oy + dy) * SL_B + ox + dx) * SL_PS;
          const float* wp = s.w4d + (dy * 3 + dx) * 32;
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 a = *reinterpret_cast<const float4*>(ap + c4 * 4), u = *reinterpret_cast<const float4*>(wp + c4 * 4);
            acc = fmaf(a.x, u.x, acc); acc = fmaf(a.y, u.y, acc); acc = fmaf(a.z, u.z, acc); acc = fmaf(a.w, u.w, acc);
          }
        }
      const float dv = sigmoid_acc(acc);
      s.diff[tid] = dv;
      const int gy = y0 + oy, gx = x0 + ox;
      if (gy < H && gx < W) diff[(long)b * HW + (long)gy * W + gx] = dv;
    }
    // ---- gate branch
    sel_conv_in<SL_B>(s, s.w0[1], s.t1, 2, y0 - 1, x0 - 1, H, W, tid);
    __syncthreads();                                // g0 complete, d4 done with t2, g2 weights in place
    sel_conv32<SL_B, 2>(s.t1, s.w2, s.b2[1], s.t2, y0, x0, H, W, tid);
    __syncthreads();
    if (tid < SL_T * SL_T) {
      const int oy = tid / SL_T, ox = tid % SL_T;
      const int gy = y0 + oy, gx = x0 + ox;
      float r4[4] = {s.w4g[SL_W4G], s.w4g[SL_W4G + 1], s.w4g[SL_W4G + 2], s.w4g[SL_W4G + 3]};
      const float* ap = s.t2 + tid * SL_PS;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 a = *reinterpret_cast<const float4*>(ap + c4 * 4);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 u = *reinterpret_cast<const float4*>(s.w4g + (c4 * 4 + k) * 4);
          r4[0] = fmaf(av[k], u.x, r4[0]); r4[1] = fmaf(av[k], u.y, r4[1]); r4[2] = fmaf(av[k], u.z, r4[2]); r4[3] = fmaf(av[k], u.w, r4[3]);
        }
      }
      if (gy < H && gx < W) {
        const long gp = (long)b * HW + (long)gy * W + gx;
        *reinterpret_cast<float4*>(graw + gp * 4) = make_float4(r4[0], r4[1], r4[2], r4[3]);
        // gate normalisation exactly as k_gate_finalize (lr_tokens.cu)
        const float thr = 0.7f - 0.5f * s.diff[tid];
        const float g0 = sigmoid_acc(T * (r4[0] - thr)), g1 = sigmoid_acc(T * (r4[1] - thr));
        const float g2 = sigmoid_acc(T * (r4[2] - thr)), g3 = sigmoid_acc(T * (r4[3] - thr));
        const float sm = fmaxf(((g0 + g1) + g2) + g3 + 1e-8f, 0.3f);
        float* o = gates + (long)b * 4 * HW + (long)gy * W + gx;
        o[0] = g0 / sm; o[HW] = g1 / sm; o[2 * HW] = g2 / sm; o[3 * HW] = g3 / sm;
      }
    }
  }
}
}  // namespace

extern "C" size_t ffsr_selector_blob_floats(void) { return (size_t)SL_BLOB; }

// routing [B][3][H][W] fp32 -> diff [B][1][H][W], graw [B][H][W][4] (gate_net logits), gates [B][4][H][W]
extern "C" int ffsr_selector_fused(const float* routing, int B, int H, int W, const float* blob, const float* temperature, float* diff,
                                   float* graw, float* gates, cudaStream_t stream) {
  FFSR_REQUIRE(routing && blob && temperature && diff && graw && gates && B > 0 && H > 0 && W > 0, FFSR_ERR_ARG, "selector_fused: bad argument");
  FFSR_REQUIRE(((uintptr_t)blob % 16) == 0 && ((uintptr_t)graw % 16) == 0, FFSR_ERR_ALIGN, "selector_fused: 16-byte alignment required");
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(k_selector, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SelSmem));
  }
  const long tiles = (long)B * ((H + SL_T - 1) / SL_T) * ((W + SL_T - 1) / SL_T);
  const int grid = (int)(tiles < num_sms ? tiles : num_sms);
  k_selector<<<grid, SL_THREADS, sizeof(SelSmem), stream>>>(routing, B, H, W, blob, temperature, diff, graw, gates);
  return ffsr_check_launch("selector_fused");
}
