// placeholder until the tcgen05 implicit-GEMM kernel lands
#include "common.cuh"
#include "../../include/ffsr_b200.h"
int ffsr_conv2d_tc(const ffsr_conv_params* p, cudaStream_t stream) {
  (void)p; (void)stream;
  ffsr_set_error("conv2d: bf16 tensor-core path not built");
  return FFSR_ERR_ARG;
}
