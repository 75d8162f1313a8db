// dsr_debug.h -- plain (non-TMA) operand descriptions for the checker kernels in dsr_debug.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dsr_conv.cuh"

namespace dsr {

struct ActRef {          // padded NHWC 16-bit tensor as seen through make_act_map()
  const uint16_t* ptr;
  int bf16;              // 0 fp16, 1 bf16
  int C, Wp, Hp, step;
};
struct WgtRef {          // packed weight matrix [rows][K]
  const uint16_t* ptr;
  int bf16;
  int K;
};

int launch_conv_ref(const ConvGemmParams& p, const ActRef& a, const WgtRef& b, cudaStream_t s);
int launch_wgrad_ref(const WgradParams& p, const ActRef& dr, const ActRef& x, cudaStream_t s);

}  // namespace dsr
