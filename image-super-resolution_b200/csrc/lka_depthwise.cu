// LKA depthwise chain on channels-last tensors [N][H][W][C]:
//   n  = BN1(x) inside the image, 0 outside      (LKABlock.forward, large_kernel_attention.py:146)
//   t1 = dw5x5(n)    zero pad 2                   (LargeKernelAttention.forward :98)
//   t2 = dw1x21(t1)  zero pad 10 along W          (:99)
//   t3 = dw21x1(t2)  zero pad 10 along H          (:100)
// Every conv zero-pads ITS OWN input, so intermediates are zero outside the image
// (three separately padded convs != one 25x25 conv at the borders; SURVEY §7 hard part 5).
// HBM/L1-bound: each thread owns one channel and a short run of outputs along the conv
// axis, keeps the sliding window in registers, and warps read 32 consecutive channels.
#include "common.cuh"

namespace {
constexpr int R5 = 4;    // 4x4 output patch per thread for the 5x5 (8x8 input window in registers)
constexpr int R21 = 16;  // outputs per thread along the conv axis for the 21-tap passes
}

// BN is folded to y = x*k + d (eval: running stats; train: batch stats computed upstream).
template <typename TI>
__global__ void __launch_bounds__(256) k_lka_dw5(const TI* __restrict__ x, int H, int W, int C,
                                                 const float* __restrict__ bn_k, const float* __restrict__ bn_d,
                                                 const float* __restrict__ w5, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int runs_x = (W + R5 - 1) / R5, runs_y = (H + R5 - 1) / R5;
  const int idx = blockIdx.y * blockDim.y + threadIdx.y;      // over runs_y * runs_x
  const int y0 = (idx / runs_x) * R5, x0 = (idx % runs_x) * R5;
  const int n = blockIdx.z;
  if (c >= C || idx >= runs_y * runs_x) return;
  float w[25];
#pragma unroll
  for (int i = 0; i < 25; ++i) w[i] = w5[c * 25 + i];
  const float k = bn_k[c], d = bn_d[c];
  float acc[R5][R5];
#pragma unroll
  for (int a = 0; a < R5; ++a)
#pragma unroll
    for (int b = 0; b < R5; ++b) acc[a][b] = 0.f;
  const TI* img = x + (long)n * H * W * C;
  if (y0 >= 2 && y0 + R5 + 2 <= H && x0 >= 2 && x0 + R5 + 2 <= W) {
    // interior patch: no bounds tests, 32-bit strided offsets (the kernel is issue-bound, not HBM-bound)
    const TI* p = img + ((long)(y0 - 2) * W + (x0 - 2)) * C + c;
    const int rs = W * C;
#pragma unroll
    for (int iy = 0; iy < R5 + 4; ++iy) {
      float v[R5 + 4];
#pragma unroll
      for (int i = 0; i < R5 + 4; ++i) v[i] = fmaf(to_f32<TI>(p[iy * rs + i * C]), k, d);
#pragma unroll
      for (int a = 0; a < R5; ++a) {
        const int dy = iy - a;
        if (dy < 0 || dy > 4) continue;
#pragma unroll
        for (int b = 0; b < R5; ++b)
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) acc[a][b] = fmaf(w[dy * 5 + dx], v[b + dx], acc[a][b]);
      }
    }
    float* o = out + (((long)n * H + y0) * W + x0) * C + c;
#pragma unroll
    for (int a = 0; a < R5; ++a)
#pragma unroll
      for (int b = 0; b < R5; ++b) o[a * rs + b * C] = acc[a][b];
    return;
  }
#pragma unroll
  for (int iy = 0; iy < R5 + 4; ++iy) {
    const int yy = y0 + iy - 2;
    if (yy < 0 || yy >= H) continue;                        // zero padding of n = BN1(x)
    float v[R5 + 4];
#pragma unroll
    for (int i = 0; i < R5 + 4; ++i) {
      const int xx = x0 + i - 2;
      v[i] = (xx >= 0 && xx < W) ? fmaf(to_f32<TI>(img[((long)yy * W + xx) * C + c]), k, d) : 0.f;
    }
#pragma unroll
    for (int a = 0; a < R5; ++a) {
      const int dy = iy - a;                                // tap row that maps input row iy to output row a
      if (dy < 0 || dy > 4) continue;
#pragma unroll
      for (int b = 0; b < R5; ++b)
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) acc[a][b] = fmaf(w[dy * 5 + dx], v[b + dx], acc[a][b]);
    }
  }
#pragma unroll
  for (int a = 0; a < R5; ++a)
#pragma unroll
    for (int b = 0; b < R5; ++b)
      if (y0 + a < H && x0 + b < W) out[(((long)n * H + y0 + a) * W + x0 + b) * C + c] = acc[a][b];
}

// AXIS = 0: taps along W (1x21); AXIS = 1: taps along H (21x1)
template <int AXIS, typename TO>
__global__ void __launch_bounds__(256) k_lka_dw21(const float* __restrict__ in, int H, int W, int C,
                                                  const float* __restrict__ w21, TO* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.z;
  int y, x;
  if (AXIS == 0) {
    const int runs = (W + R21 - 1) / R21;
    const int idx = blockIdx.y * blockDim.y + threadIdx.y;    // over H * runs
    y = idx / runs;
    x = (idx % runs) * R21;
    if (y >= H) return;
  } else {
    const int runs = (H + R21 - 1) / R21;
    const int idx = blockIdx.y * blockDim.y + threadIdx.y;    // over runs * W
    y = (idx / W) * R21;
    x = idx % W;
    if (idx >= runs * W) return;
  }
  if (c >= C) return;
  float w[21];
#pragma unroll
  for (int i = 0; i < 21; ++i) w[i] = w21[c * 21 + i];
  const float* img = in + (long)n * H * W * C;
  TO* oimg = out + (long)n * H * W * C;
  float v[R21 + 20];
  // The kernel is instruction-issue bound (ncu: 74 % issue slots busy at 25 % of HBM), so runs that do not touch
  // the border take a path without per-load bounds tests and with 32-bit strided offsets.
  const int stride = (AXIS == 0) ? C : W * C;
  const bool interior = (AXIS == 0) ? (x >= 10 && x + R21 + 10 <= W) : (y >= 10 && y + R21 + 10 <= H);
  if (interior) {
    const float* p = img + ((AXIS == 0) ? ((long)y * W + (x - 10)) : ((long)(y - 10) * W + x)) * C + c;
#pragma unroll
    for (int i = 0; i < R21 + 20; ++i) v[i] = p[i * stride];
    TO* o = oimg + ((long)y * W + x) * C + c;
#pragma unroll
    for (int r = 0; r < R21; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < 21; ++t) acc = fmaf(w[t], v[r + t], acc);
      o[r * stride] = from_f32<TO>(acc);
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < R21 + 20; ++i) {
    if (AXIS == 0) {
      const int xx = x + i - 10;
      v[i] = (xx >= 0 && xx < W) ? img[((long)y * W + xx) * C + c] : 0.f;
    } else {
      const int yy = y + i - 10;
      v[i] = (yy >= 0 && yy < H) ? img[((long)yy * W + x) * C + c] : 0.f;
    }
  }
#pragma unroll
  for (int r = 0; r < R21; ++r) {
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 21; ++t) acc = fmaf(w[t], v[r + t], acc);
    if (AXIS == 0) {
      if (x + r < W) oimg[((long)y * W + x + r) * C + c] = from_f32<TO>(acc);
    } else {
      if (y + r < H) oimg[((long)(y + r) * W + x) * C + c] = from_f32<TO>(acc);
    }
  }
}

// x: [N][H][W][C] -> out: [N][H][W][C]; tmp1/tmp2: same-size scratch (tmp2 may alias out? no: distinct)
static int lka_depthwise_impl(const void* x, int x_dtype, int N, int H, int W, int C, const float* bn_k, const float* bn_d,
                              const float* w5, const float* wh, const float* wv, float* tmp1, float* tmp2, void* out,
                              int out_dtype, cudaStream_t stream);

extern "C" int ffsr_lka_depthwise(const float* x, int N, int H, int W, int C, const float* bn_k, const float* bn_d,
                                  const float* w5, const float* wh, const float* wv, float* tmp1, float* tmp2,
                                  void* out, int out_dtype, cudaStream_t stream) {
  return lka_depthwise_impl(x, 0, N, H, W, C, bn_k, bn_d, w5, wh, wv, tmp1, tmp2, out, out_dtype, stream);
}
// same chain with a bf16 input tensor (bf16 mode: the Phase-4 residual stream is stored as bf16)
extern "C" int ffsr_lka_depthwise_in(const void* x, int x_dtype, int N, int H, int W, int C, const float* bn_k,
                                     const float* bn_d, const float* w5, const float* wh, const float* wv, float* tmp1,
                                     float* tmp2, void* out, int out_dtype, cudaStream_t stream) {
  return lka_depthwise_impl(x, x_dtype, N, H, W, C, bn_k, bn_d, w5, wh, wv, tmp1, tmp2, out, out_dtype, stream);
}

static int lka_depthwise_impl(const void* x, int x_dtype, int N, int H, int W, int C, const float* bn_k, const float* bn_d,
                              const float* w5, const float* wh, const float* wv, float* tmp1, float* tmp2, void* out,
                              int out_dtype, cudaStream_t stream) {
  FFSR_REQUIRE(x && bn_k && bn_d && w5 && wh && wv && tmp1 && tmp2 && out, FFSR_ERR_ARG, "lka_depthwise: null pointer");
  FFSR_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 32 == 0, FFSR_ERR_ARG, "lka_depthwise: C must be a multiple of 32");
  FFSR_REQUIRE(N <= 65535 && (long)H * W / 4 < 65535L * 4, FFSR_ERR_ARG, "lka_depthwise: grid too large");
  const int cx = C < 64 ? C : 64;
  const int ty = 256 / cx;
  {
    dim3 block(cx, ty);
    dim3 grid(C / cx, ceil_div((long)ceil_div(H, R5) * ceil_div(W, R5), ty), N);
    if (x_dtype == 1) k_lka_dw5<__nv_bfloat16><<<grid, block, 0, stream>>>((const __nv_bfloat16*)x, H, W, C, bn_k, bn_d, w5, tmp1);
    else k_lka_dw5<float><<<grid, block, 0, stream>>>((const float*)x, H, W, C, bn_k, bn_d, w5, tmp1);
    int rc = ffsr_check_launch("lka_dw5");
    if (rc) return rc;
  }
  {
    dim3 block(cx, ty);
    const long items = (long)H * ceil_div(W, R21);
    dim3 grid(C / cx, ceil_div(items, ty), N);
    k_lka_dw21<0, float><<<grid, block, 0, stream>>>(tmp1, H, W, C, wh, tmp2);
    int rc = ffsr_check_launch("lka_dw21_h");
    if (rc) return rc;
  }
  {
    dim3 block(cx, ty);
    const long items = (long)ceil_div(H, R21) * W;
    dim3 grid(C / cx, ceil_div(items, ty), N);
    if (out_dtype == 1)
      k_lka_dw21<1, __nv_bfloat16><<<grid, block, 0, stream>>>(tmp2, H, W, C, wv, (__nv_bfloat16*)out);
    else
      k_lka_dw21<1, float><<<grid, block, 0, stream>>>(tmp2, H, W, C, wv, (float*)out);
    return ffsr_check_launch("lka_dw21_v");
  }
}

// ------------------------------------------------------------------------------------
// Train-mode access to the individual stages (the backward pass runs them with flipped taps:
// the input gradient of a zero-padded stride-1 depthwise conv is the same conv with reversed
// taps), and the depthwise weight gradients  dw[c][tap] = sum_{n,y,x} in[n,y+dy,x+dx,c] * g[n,y,x,c].
// kind: 0 = 5x5 (pad 2), 1 = 1x21 along W (pad 10), 2 = 21x1 along H (pad 10).
// ------------------------------------------------------------------------------------
extern "C" int ffsr_dwconv_stage(const float* in, int N, int H, int W, int C, int kind, const float* w,
                                 const float* bn_k, const float* bn_d, float* out, cudaStream_t stream) {
  FFSR_REQUIRE(in && w && out, FFSR_ERR_ARG, "dwconv_stage: null pointer");
  FFSR_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C % 32 == 0, FFSR_ERR_ARG, "dwconv_stage: C must be a multiple of 32");
  FFSR_REQUIRE(N <= 65535 && (long)H * W / 4 < 65535L * 4, FFSR_ERR_ARG, "dwconv_stage: grid too large");
  FFSR_REQUIRE(kind >= 0 && kind <= 2, FFSR_ERR_ARG, "dwconv_stage: kind must be 0, 1 or 2");
  FFSR_REQUIRE(kind != 0 || (bn_k && bn_d), FFSR_ERR_ARG, "dwconv_stage: the 5x5 stage needs the input affine (ones/zeros for none)");
  const int cx = C < 64 ? C : 64;
  const int ty = 256 / cx;
  dim3 block(cx, ty);
  if (kind == 0) {
    dim3 grid(C / cx, ceil_div((long)ceil_div(H, R5) * ceil_div(W, R5), ty), N);
    k_lka_dw5<float><<<grid, block, 0, stream>>>(in, H, W, C, bn_k, bn_d, w, out);
  } else if (kind == 1) {
    dim3 grid(C / cx, ceil_div((long)H * ceil_div(W, R21), ty), N);
    k_lka_dw21<0, float><<<grid, block, 0, stream>>>(in, H, W, C, w, out);
  } else {
    dim3 grid(C / cx, ceil_div((long)ceil_div(H, R21) * W, ty), N);
    k_lka_dw21<1, float><<<grid, block, 0, stream>>>(in, H, W, C, w, out);
  }
  return ffsr_check_launch("dwconv_stage");
}

namespace {
constexpr int WG_R = 8;    // outputs per thread along the conv axis in the 21-tap weight-gradient kernel

// AXIS 0: taps along W, AXIS 1: taps along H.  blockDim = (channels, items); each thread owns one
// channel and WG_R consecutive outputs; per-block partial sums in shared memory, then global atomics.
template <int AXIS>
__global__ void __launch_bounds__(256) k_dw21_wgrad(const float* __restrict__ in, const float* __restrict__ g, int H,
                                                    int W, int C, long items, float* __restrict__ dw) {
  extern __shared__ float sdw[];            // [blockDim.x][21]
  const int tl = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tl; i < blockDim.x * 21; i += blockDim.x * blockDim.y) sdw[i] = 0.f;
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.z;
  const long idx = (long)blockIdx.y * blockDim.y + threadIdx.y;
  if (c < C && idx < items) {
    int y, x;
    if (AXIS == 0) {
      const int runs = (W + WG_R - 1) / WG_R;
      y = (int)(idx / runs);
      x = (int)(idx % runs) * WG_R;
    } else {
      y = (int)(idx / W) * WG_R;
      x = (int)(idx % W);
    }
    const float* img = in + (long)n * H * W * C;
    const float* gim = g + (long)n * H * W * C;
    float v[WG_R + 20], gr[WG_R];
#pragma unroll
    for (int i = 0; i < WG_R + 20; ++i) {
      if (AXIS == 0) {
        const int xx = x + i - 10;
        v[i] = (xx >= 0 && xx < W) ? img[((long)y * W + xx) * C + c] : 0.f;
      } else {
        const int yy = y + i - 10;
        v[i] = (yy >= 0 && yy < H) ? img[((long)yy * W + x) * C + c] : 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < WG_R; ++r) {
      if (AXIS == 0) gr[r] = (x + r < W) ? gim[((long)y * W + x + r) * C + c] : 0.f;
      else gr[r] = (y + r < H) ? gim[((long)(y + r) * W + x) * C + c] : 0.f;
    }
#pragma unroll
    for (int t = 0; t < 21; ++t) {
      float a = 0.f;
#pragma unroll
      for (int r = 0; r < WG_R; ++r) a = fmaf(v[r + t], gr[r], a);
      atomicAdd(&sdw[threadIdx.x * 21 + t], a);
    }
  }
  __syncthreads();
  for (int i = tl; i < blockDim.x * 21; i += blockDim.x * blockDim.y) {
    const int cc = blockIdx.x * blockDim.x + i / 21;
    if (cc < C) atomicAdd(dw + (long)cc * 21 + i % 21, sdw[i]);
  }
}

__global__ void __launch_bounds__(256) k_dw5_wgrad(const float* __restrict__ in, const float* __restrict__ g, int H,
                                                   int W, int C, long items, float* __restrict__ dw) {
  extern __shared__ float sdw[];            // [blockDim.x][25]
  const int tl = threadIdx.y * blockDim.x + threadIdx.x;
  for (int i = tl; i < blockDim.x * 25; i += blockDim.x * blockDim.y) sdw[i] = 0.f;
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.z;
  const long idx = (long)blockIdx.y * blockDim.y + threadIdx.y;
  if (c < C && idx < items) {
    const int runs_x = (W + R5 - 1) / R5;
    const int y0 = (int)(idx / runs_x) * R5, x0 = (int)(idx % runs_x) * R5;
    const float* img = in + (long)n * H * W * C;
    const float* gim = g + (long)n * H * W * C;
    float gr[R5][R5];
#pragma unroll
    for (int a = 0; a < R5; ++a)
#pragma unroll
      for (int b = 0; b < R5; ++b)
        gr[a][b] = (y0 + a < H && x0 + b < W) ? gim[((long)(y0 + a) * W + x0 + b) * C + c] : 0.f;
    float acc[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) acc[i] = 0.f;
#pragma unroll
    for (int iy = 0; iy < R5 + 4; ++iy) {
      const int yy = y0 + iy - 2;
      float v[R5 + 4];
#pragma unroll
      for (int i = 0; i < R5 + 4; ++i) {
        const int xx = x0 + i - 2;
        v[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? img[((long)yy * W + xx) * C + c] : 0.f;
      }
#pragma unroll
      for (int a = 0; a < R5; ++a) {
        const int dy = iy - a;
        if (dy < 0 || dy > 4) continue;
#pragma unroll
        for (int b = 0; b < R5; ++b)
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) acc[dy * 5 + dx] = fmaf(v[b + dx], gr[a][b], acc[dy * 5 + dx]);
      }
    }
#pragma unroll
    for (int i = 0; i < 25; ++i) atomicAdd(&sdw[threadIdx.x * 25 + i], acc[i]);
  }
  __syncthreads();
  for (int i = tl; i < blockDim.x * 25; i += blockDim.x * blockDim.y) {
    const int cc = blockIdx.x * blockDim.x + i / 25;
    if (cc < C) atomicAdd(dw + (long)cc * 25 + i % 25, sdw[i]);
  }
}
}  // namespace

// dw: [C][taps] fp32, ACCUMULATED into (caller zeroes)
extern "C" int ffsr_dwconv_wgrad(const float* in, const float* g, int N, int H, int W, int C, int kind, float* dw,
                                 cudaStream_t stream) {
  FFSR_REQUIRE(in && g && dw, FFSR_ERR_ARG, "dwconv_wgrad: null pointer");
  FFSR_REQUIRE(N > 0 && N <= 65535 && H > 0 && W > 0 && C > 0 && C % 32 == 0, FFSR_ERR_ARG, "dwconv_wgrad: bad shape");
  FFSR_REQUIRE(kind >= 0 && kind <= 2, FFSR_ERR_ARG, "dwconv_wgrad: kind must be 0, 1 or 2");
  const int cx = 32, ty = 8;
  dim3 block(cx, ty);
  if (kind == 0) {
    const long items = (long)ceil_div(H, R5) * ceil_div(W, R5);
    dim3 grid(C / cx, ceil_div(items, ty), N);
    k_dw5_wgrad<<<grid, block, cx * 25 * sizeof(float), stream>>>(in, g, H, W, C, items, dw);
  } else if (kind == 1) {
    const long items = (long)H * ceil_div(W, WG_R);
    dim3 grid(C / cx, ceil_div(items, ty), N);
    k_dw21_wgrad<0><<<grid, block, cx * 21 * sizeof(float), stream>>>(in, g, H, W, C, items, dw);
  } else {
    const long items = (long)ceil_div(H, WG_R) * W;
    dim3 grid(C / cx, ceil_div(items, ty), N);
    k_dw21_wgrad<1><<<grid, block, cx * 21 * sizeof(float), stream>>>(in, g, H, W, C, items, dw);
  }
  return ffsr_check_launch("dwconv_wgrad");
}
