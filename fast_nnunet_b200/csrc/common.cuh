// Shared helpers for libfnnu (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fnnu.h"

namespace fnnu {

void set_error(const char* fmt, ...);

#define FNNU_CHECK_ARG(cond, ...)                 \
  do {                                            \
    if (!(cond)) {                                \
      fnnu::set_error(__VA_ARGS__);               \
      return FNNU_E_INVALID;                      \
    }                                             \
  } while (0)

#define FNNU_CUDA(expr)                                                              \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      fnnu::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return FNNU_E_CUDA;                                                            \
    }                                                                                \
  } while (0)

#define FNNU_LAUNCH_CHECK()                                                          \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      fnnu::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return FNNU_E_CUDA;                                                            \
    }                                                                                \
  } while (0)

// Per-channel pending transform of an activation buffer: y = lrelu_slope(x * scale + shift).
// meta = {gamma, beta, slope, eps}; eps < 0 marks a channel that is stored ready to use.
struct ChanMeta {
  float gamma, beta, slope, eps;
};

// View of `c` channels of a channels-last fp16 buffer [batch][d0][d1][d2][cs].
struct BufView {
  __half* ptr;          // first channel of the view (already offset by coff)
  int d[3];
  int cs;               // channel stride of the underlying buffer
  const double* stats;  // [batch][cs][2] of the underlying buffer, offset to the view's first channel
  const ChanMeta* meta; // [cs], offset likewise
  int stat_stride;      // = cs (doubles pairs per sample)
};

__host__ __device__ inline size_t vox(const int d[3]) { return (size_t)d[0] * d[1] * d[2]; }

// scale/shift of one (sample, channel) from the InstanceNorm sums (biased variance, eps inside sqrt).
__device__ __forceinline__ void xform_from_stats(const double* st, const ChanMeta m, double inv_count,
                                                 float& scale, float& shift) {
  if (m.eps < 0.f) {
    scale = 1.f;
    shift = 0.f;
    return;
  }
  double mean = st[0] * inv_count;
  double var = st[1] * inv_count - mean * mean;
  if (var < 0.0) var = 0.0;
  float rstd = (float)(1.0 / sqrt(var + (double)m.eps));
  scale = m.gamma * rstd;
  shift = m.beta - (float)mean * scale;
}

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

}  // namespace fnnu
