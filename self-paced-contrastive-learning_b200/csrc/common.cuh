// Shared host/device helpers for the spcl C-ABI library.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/spcl.h"

namespace spcl {

void set_last_cuda_error(cudaError_t e, const char* where);

#define SPCL_CUDA_TRY(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) {                                  \
      ::spcl::set_last_cuda_error(_e, #expr);                 \
      return SPCL_ERR_CUDA;                                   \
    }                                                         \
  } while (0)

#define SPCL_LAUNCH_CHECK(name)                               \
  do {                                                        \
    cudaError_t _e = cudaGetLastError();                      \
    if (_e != cudaSuccess) {                                  \
      ::spcl::set_last_cuda_error(_e, name);                  \
      return SPCL_ERR_CUDA;                                   \
    }                                                         \
  } while (0)

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// self-paced weight on a positive pair with loss l = -LLH (contrast_loss3.py:207-214)
__device__ __forceinline__ float sp_weight(float l, float gamma, float inv_gamma, int mode) {
  if (mode == SPCL_MODE_HARD) return l <= gamma ? 1.f : 0.f;
  if (mode == SPCL_MODE_SOFT) return fmaxf(1.f - l * inv_gamma, 0.f);
  return 1.f;
}

// 1 / gamma for the soft rule.  gamma == 0 is legal in the reference (PScheduler's default begin_value, infonce.py:34-53):
// 1 - l / 0 = -inf for every l > 0, so every positive weight is 0 and the loss is 0; FLT_MAX gives the same weights
// without producing inf * 0.
inline float inv_gamma_of(float gamma) { return gamma > 0.f ? 1.f / gamma : 3.402823466e+38f; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace spcl
