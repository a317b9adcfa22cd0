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

// Exact-erf GELU, x * Phi(x), for the GEMM epilogues:  gelu(x) = relu(x) - |x| * Q(|x|),  Q(a) = 0.5 erfc(a / sqrt 2).
// log2 Q is smooth, so Q(a) = exp2(P8(a)) with a degree-8 polynomial (weighted least-squares fit on [0, 6.5]; beyond
// that |x| Q < 1e-9): max |error| 2.5e-7 over all x (the fp32 rounding of the result), 1e-5 relative in the far
// negative tail.  8 FMA + ONE MUFU (ex2) instead of rcp + ex2 + 12 FMA-pipe instructions: in the epilogue of the
// K-short ConvNeXt layers the GELU was MUFU-bound (two MUFU per element = 16 issue cycles per warp).
__device__ __forceinline__ float gelu_exact(float x) {
  const float a = fminf(fabsf(x), 6.5f);
  float p = -8.853008922e-07f;
  p = fmaf(p, a, 1.435392369e-05f);
  p = fmaf(p, a, -5.635936031e-05f);
  p = fmaf(p, a, -4.886629758e-04f);
  p = fmaf(p, a, 7.600210607e-03f);
  p = fmaf(p, a, -5.296006426e-02f);
  p = fmaf(p, a, -4.589921236e-01f);
  p = fmaf(p, a, -1.151152968e+00f);
  p = fmaf(p, a, -9.999963641e-01f);
  float q;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(p));
  return fmaxf(x, 0.f) - a * q;   // (a, not |x|: keeps gelu(+-inf) = +inf / -0 instead of inf - inf)
}
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
