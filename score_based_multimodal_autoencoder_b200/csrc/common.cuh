// Shared host/device helpers for the libsbmae_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace sbm {

// last-error string returned by sbm_last_error(); set by every failing entry point
void set_error(const char* fmt, ...);

#define SBM_CHECK_ARG(cond, ...)  \
  do {                            \
    if (!(cond)) {                \
      ::sbm::set_error(__VA_ARGS__); \
      return 1;                   \
    }                             \
  } while (0)

#define SBM_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::sbm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, i.e. fp32-level): one MUFU.RCP + one MUFU.EX2 + 7 FMA,
// branch-free -- about half the issue slots of erff(), which matters in the GEMM epilogue.
__device__ __forceinline__ float erf_as(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = exp2f(-1.4426950408889634f * ax * ax);
  return copysignf(1.0f - p * e, x);
}
__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erf_as(x * 0.70710678118654752440f)); }
// d/dx of exact-erf GELU
__device__ __forceinline__ float gelu_exact_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float silu_grad(float x) {
  const float s = 1.0f / (1.0f + __expf(-x));
  return s * (1.0f + x * (1.0f - s));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of a float; `red` is >= 32 floats of shared memory. Result valid in all threads.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? red[lane] : 0.0f;
  r = warp_sum(r);
  return r;
}

}  // namespace sbm
