// Elementwise glue of the residual autoencoders either side of the score-model path (h_vae_model_copy.py:9-39):
//   y = LeakyReLU_slope(x)  followed by  AvgPool2d(R) | nn.Upsample(scale_factor=R, nearest) | nothing,
// over channels-last activations.  The convolutions (BatchNorm folded into their weights, residual added in the GEMM
// epilogue) run on sbm_conv_igemm; this kernel is the `self.sf(x + xhat)` + `down_pool` / `up_pool` tail of RBlock,
// the LeakyReLU + AvgPool2d(2) of ResEncoder.ch_enc and (slope 0) the ReLU after ResAE.z_lin.  HBM-bound: 4 B read +
// 2 B / R^2 (pool) or 2 B * R^2 (up-sample) written per element.
#include <algorithm>
#include <atomic>

#include "../../include/sbmae_b200.h"
#include "common.cuh"

namespace sbm {
extern std::atomic<unsigned long long> g_launches;

template <typename TIn>
__device__ __forceinline__ void load8(const TIn* p, float (&v)[8]) {
  if constexpr (sizeof(TIn) == 4) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[2 * k] = __low2float(h[k]);
      v[2 * k + 1] = __high2float(h[k]);
    }
  }
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
    w[k] = *reinterpret_cast<const uint32_t*>(&t);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int ACT>
__device__ __forceinline__ float actf(float v, float slope) {
  if constexpr (ACT == 1) return gelu_exact(v);   // RBlockN: nn.GELU()
  return v > 0.f ? v : slope * v;                 // LeakyReLU(slope); slope 0 = ReLU
}

// bilinear up-sampling by R (nn.Upsample(scale_factor=R, mode='bilinear'), align_corners=False) of act(x): one thread =
// one channel octet of one OUTPUT pixel; source coordinate s = max((o + 0.5) / R - 0.5, 0), neighbours clamped at the edge
template <typename TIn, int ACT>
__global__ void __launch_bounds__(256)
act_bilinear_kernel(const TIn* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out, int64_t ldo, int B, int H,
                    int W, int C, float slope, int R) {
  const int oct = (C + 7) >> 3;
  const int OH = H * R, OW = W * R;
  const float invR = 1.f / (float)R;
  const int64_t total = (int64_t)B * OH * OW * oct;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(idx % oct);
    const int64_t pix = idx / oct;
    const int ox = (int)(pix % OW);
    const int oy = (int)((pix / OW) % OH);
    const int b = (int)(pix / ((int64_t)OW * OH));
    const float sy = fmaxf(((float)oy + 0.5f) * invR - 0.5f, 0.f), sx = fmaxf(((float)ox + 0.5f) * invR - 0.5f, 0.f);
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
    const float wy = sy - (float)y0, wx = sx - (float)x0;
    const int c0 = o * 8;
    const TIn* base = x + (int64_t)b * H * W * ldx + c0;
    float v00[8], v01[8], v10[8], v11[8], r[8];
    load8(base + ((int64_t)y0 * W + x0) * ldx, v00);
    load8(base + ((int64_t)y0 * W + x1) * ldx, v01);
    load8(base + ((int64_t)y1 * W + x0) * ldx, v10);
    load8(base + ((int64_t)y1 * W + x1) * ldx, v11);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float top = (1.f - wx) * actf<ACT>(v00[k], slope) + wx * actf<ACT>(v01[k], slope);
      const float bot = (1.f - wx) * actf<ACT>(v10[k], slope) + wx * actf<ACT>(v11[k], slope);
      r[k] = (c0 + k < C) ? (1.f - wy) * top + wy * bot : 0.f;
    }
    *reinterpret_cast<uint4*>(out + (((int64_t)b * OH + oy) * OW + ox) * ldo + c0) = pack8(r);
  }
}

// mode 0: same size; 1: average pooling by R; 2: nearest up-sampling by R.  One thread = one channel octet of one
// OUTPUT pixel (modes 0, 1) or of one INPUT pixel (mode 2).  out_bf16: channels-last [.., ldo]; out_nchw: fp32
// [B][C][OH][OW] (mode 0 only: feeds the im2col of the 5x5 output convolution).
template <typename TIn, int ACT>
__global__ void __launch_bounds__(256)
lrelu_resample_kernel(const TIn* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out, int64_t ldo,
                      float* __restrict__ out_nchw, int B, int H, int W, int C, float slope, int mode, int R) {
  const int oct = (C + 7) >> 3;
  const int PH = mode == 1 ? H / R : H, PW = mode == 1 ? W / R : W;   // pixel grid the threads walk
  const int64_t total = (int64_t)B * PH * PW * oct;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(idx % oct);
    const int64_t pix = idx / oct;
    const int pw = (int)(pix % PW);
    const int ph = (int)((pix / PW) % PH);
    const int b = (int)(pix / ((int64_t)PW * PH));
    const int c0 = o * 8;
    float acc[8];
    if (mode == 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
      for (int i = 0; i < R; ++i)
        for (int j = 0; j < R; ++j) {
          float v[8];
          load8(x + (((int64_t)b * H + ph * R + i) * W + pw * R + j) * ldx + c0, v);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += actf<ACT>(v[k], slope);
        }
      const float inv = 1.f / (float)(R * R);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] *= inv;
    } else {
      load8(x + (((int64_t)b * H + ph) * W + pw) * ldx + c0, acc);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = actf<ACT>(acc[k], slope);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (c0 + k >= C) acc[k] = 0.f;   // keep the channel padding zero for the next GEMM
    if (mode == 2) {
      const uint4 pk = pack8(acc);
      const int OW = W * R;
      for (int i = 0; i < R; ++i)
        for (int j = 0; j < R; ++j)
          *reinterpret_cast<uint4*>(out + (((int64_t)b * H * R + ph * R + i) * OW + pw * R + j) * ldo + c0) = pk;
    } else {
      if (out != nullptr) *reinterpret_cast<uint4*>(out + (((int64_t)b * PH + ph) * PW + pw) * ldo + c0) = pack8(acc);
      if (out_nchw != nullptr) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (c0 + k < C) out_nchw[(((int64_t)b * C + c0 + k) * PH + ph) * PW + pw] = acc[k];
      }
    }
  }
}

}  // namespace sbm

using namespace sbm;

extern "C" int sbm_act_resample(const void* x, int32_t in_dtype, int64_t ldx, void* out_bf16, int64_t ldo,
                                float* out_nchw_f32, int32_t B, int32_t H, int32_t W, int32_t C, int32_t act, float slope,
                                int32_t mode, int32_t rate, void* stream) {
  SBM_CHECK_ARG(x && (out_bf16 || out_nchw_f32) && B > 0 && H > 0 && W > 0 && C > 0, "sbm_act_resample: bad args");
  SBM_CHECK_ARG(act == 0 || act == 1, "sbm_act_resample: act must be 0 (LeakyReLU) or 1 (GELU)");
  SBM_CHECK_ARG(mode >= 0 && mode <= 3 && (mode == 0 || rate >= 1), "sbm_act_resample: bad mode / rate");
  SBM_CHECK_ARG(mode != 1 || (H % rate == 0 && W % rate == 0), "sbm_act_resample: %dx%d not divisible by %d", H, W, rate);
  SBM_CHECK_ARG(mode == 0 || out_nchw_f32 == nullptr, "sbm_act_resample: NCHW output only without resampling");
  SBM_CHECK_ARG(mode < 2 || out_bf16 != nullptr, "sbm_act_resample: up-sampling writes the bf16 output");
  const int esz = in_dtype == SBM_F32 ? 4 : 2;
  SBM_CHECK_ARG(in_dtype == SBM_F32 || in_dtype == SBM_BF16, "sbm_act_resample: dtype");
  SBM_CHECK_ARG(ldx % 8 == 0 && ldx >= ((C + 7) / 8) * 8 && (reinterpret_cast<uintptr_t>(x) % (4 * esz)) == 0,
                "sbm_act_resample: input rows must be 16-byte aligned and cover C rounded up to 8");
  SBM_CHECK_ARG(out_bf16 == nullptr || (ldo % 8 == 0 && ldo >= ((C + 7) / 8) * 8 &&
                                        (reinterpret_cast<uintptr_t>(out_bf16) & 15) == 0),
                "sbm_act_resample: output rows must be 16-byte aligned and cover C rounded up to 8");
  const int PH = mode == 1 ? H / rate : (mode == 3 ? H * rate : H), PW = mode == 1 ? W / rate : (mode == 3 ? W * rate : W);
  const int64_t total = (int64_t)B * PH * PW * ((C + 7) / 8);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16));
  cudaStream_t st = (cudaStream_t)stream;
#define SBM_AR(TIN, ACT)                                                                                              \
  do {                                                                                                                \
    if (mode == 3)                                                                                                    \
      act_bilinear_kernel<TIN, ACT><<<grid, 256, 0, st>>>((const TIN*)x, ldx, (__nv_bfloat16*)out_bf16, ldo, B, H, W, C, \
                                                          slope, rate);                                              \
    else                                                                                                              \
      lrelu_resample_kernel<TIN, ACT><<<grid, 256, 0, st>>>((const TIN*)x, ldx, (__nv_bfloat16*)out_bf16, ldo,          \
                                                            out_nchw_f32, B, H, W, C, slope, mode, rate);            \
  } while (0)
  if (in_dtype == SBM_F32) {
    if (act == 1) SBM_AR(float, 1);
    else SBM_AR(float, 0);
  } else {
    if (act == 1) SBM_AR(__nv_bfloat16, 1);
    else SBM_AR(__nv_bfloat16, 0);
  }
#undef SBM_AR
  SBM_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

extern "C" int sbm_lrelu_resample(const void* x, int32_t in_dtype, int64_t ldx, void* out_bf16, int64_t ldo,
                                  float* out_nchw_f32, int32_t B, int32_t H, int32_t W, int32_t C, float slope,
                                  int32_t mode, int32_t rate, void* stream) {
  SBM_CHECK_ARG(mode >= 0 && mode <= 2, "sbm_lrelu_resample: bad mode");
  return sbm_act_resample(x, in_dtype, ldx, out_bf16, ldo, out_nchw_f32, B, H, W, C, 0, slope, mode, rate, stream);
}
