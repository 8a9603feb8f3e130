// Fused optimizer step over ONE flat fp32 parameter bucket:
//   global-norm gradient clipping  torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)   train.py:344-348
//   AdamW                          torch.optim.AdamW(lr, betas, eps, weight_decay)           train.py:847-853
//   EMA shadow update              EMAModel.update                    src/utils/checkpoint_manager.py:352-359
// The reference runs these as ~600 small ATen launches per step (198 tensors x 3); here it is one
// reduction + one elementwise pass: 20 B read + 16 B written per parameter.  No host sync: the clip
// coefficient is derived on the device from the squared-norm accumulator.
#include "common.cuh"
#include "../../include/ffsr_b200.h"

namespace {
__global__ void __launch_bounds__(256) k_sumsq(const float* __restrict__ g, long n, double* __restrict__ out) {
  __shared__ double sh[8];
  double acc = 0.0;
  float f = 0.f;
  int cnt = 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    f = fmaf(g[i], g[i], f);
    if (++cnt == 16) { acc += f; f = 0.f; cnt = 0; }
  }
  acc += f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += sh[i];
    atomicAdd(out, s);
  }
}

__global__ void __launch_bounds__(256) k_adamw_ema(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v,
                                                   float* __restrict__ ema, long n, float lr, float b1, float b2,
                                                   float eps, float wd, int step, const int* __restrict__ step_dev,
                                                   const float* __restrict__ lr_dev,
                                                   const double* __restrict__ gsumsq, float grad_scale,
                                                   float max_norm, float ema_decay) {
  if (step_dev) step += *step_dev;
  if (lr_dev) lr = *lr_dev;
  const float bc1 = 1.0f - (float)pow((double)b1, (double)step);
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, (double)step));
  float clip = 1.0f;
  if (gsumsq && max_norm > 0.f) {
    const float norm = grad_scale * (float)sqrt(*gsumsq);
    clip = fminf(1.0f, max_norm / (norm + 1e-6f));
  }
  const float gs = grad_scale * clip;
  const float step_size = lr / bc1;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    pi -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (ema) ema[i] = ema_decay * ema[i] + (1.0f - ema_decay) * pi;
  }
}
}  // namespace

extern "C" int ffsr_sumsq(const float* g, long n, double* out, cudaStream_t stream) {
  FFSR_REQUIRE(g && out && n > 0, FFSR_ERR_ARG, "sumsq: bad argument");
  const int grid = (int)((n + 255) / 256 < 148L * 4 ? (n + 255) / 256 : 148L * 4);
  k_sumsq<<<grid, 256, 0, stream>>>(g, n, out);
  return ffsr_check_launch("sumsq");
}

extern "C" int ffsr_adamw_ema_step(float* p, const float* g, float* m, float* v, float* ema, long n, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, int step,
                                   const int* step_dev, const float* lr_dev, const double* gsumsq, float grad_scale,
                                   float max_norm, float ema_decay, cudaStream_t stream) {
  FFSR_REQUIRE(p && g && m && v && n > 0 && (step >= 1 || step_dev), FFSR_ERR_ARG, "adamw_ema_step: bad argument");
  const int grid = (int)((n + 255) / 256 < 148L * 8 ? (n + 255) / 256 : 148L * 8);
  k_adamw_ema<<<grid, 256, 0, stream>>>(p, g, m, v, ema, n, lr, beta1, beta2, eps, weight_decay, step, step_dev, lr_dev,
                                        gsumsq, grad_scale, max_norm, ema_decay);
  return ffsr_check_launch("adamw_ema_step");
}
