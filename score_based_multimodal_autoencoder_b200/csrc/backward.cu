// Backward kernels of the memory-bound score-net operators + fused Adam, for the DSM training step
// (loss.backward(); optimizer.step() -- train_lat_celebhq_unet_cont2.py:98-100, train_poly_unet_cont.py:272-276).
//   * column sums (bias gradients)
//   * GroupNorm backward (reduce + apply, optional GELU/SiLU chain on the normalised INPUT)
//   * depthwise 7x7 backward w.r.t. weights / bias / time condition  (input gradient = forward kernel with flipped taps)
//   * linear-attention and softmax-attention core backward
//   * activation backward, NCHW->channels-last gradient staging, fp32 accumulate
//   * multi-tensor Adam (torch.optim.Adam semantics: no amsgrad, no weight decay)
#include <atomic>

#include "../../include/sbmae_b200.h"
#include "common.cuh"

namespace sbm {
extern std::atomic<unsigned long long> g_launches;
static inline void count_launch_b() { g_launches.fetch_add(1, std::memory_order_relaxed); }

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

__device__ __forceinline__ float act_fwd(float x, int act) {
  return act == SBM_ACT_GELU ? gelu_exact(x) : (act == SBM_ACT_SILU ? silu(x) : x);
}
__device__ __forceinline__ float act_grad(float x, int act) {
  return act == SBM_ACT_GELU ? gelu_exact_grad(x) : (act == SBM_ACT_SILU ? silu_grad(x) : 1.f);
}

// ------------------------------------------------------------------------------ column sums
// out[c] += sum over rows of x[row][c]   (caller zeroes out)
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, int64_t ld, int64_t rows, int C, float* __restrict__ out) {
  // thread (cx, ry): column cx + k*32..., rows strided by 8*gridDim
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  __shared__ float red[8][33];
  for (int c0 = blockIdx.y * 32; c0 < C; c0 += gridDim.y * 32) {
    const int c = c0 + cx;
    float acc = 0.f;
    if (c < C)
      for (int64_t r = blockIdx.x * 8 + ry; r < rows; r += (int64_t)gridDim.x * 8) acc += ldf<T>(x + r * ld + c);
    red[ry][cx] = acc;
    __syncthreads();
    if (ry == 0 && c < C) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += red[k][cx];
      atomicAdd(out + c, s);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------ GroupNorm backward
// y = xh*gamma + beta, xh = (x - mean)*rstd, x = act(pre) when in_act != 0 (GroupNorm applied to an activation),
// out = out_act(y) when out_act != 0 (activation applied after the norm).
// reduce: bst[b][g] += (sum dy*gamma, sum dy*gamma*xh);  dgamma[c] += sum dy*xh; dbeta[c] += sum dy
// Thread = one channel octet (16-byte vector loads) x a strided set of pixels of ONE sample: the per-channel sums stay in
// registers over the pixel loop; shared-memory atomics happen once per thread, global atomics once per block.
template <typename T>
__device__ __forceinline__ void gld8(const T* p, float* v);
template <>
__device__ __forceinline__ void gld8<float>(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void gld8<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
  const uint4 t = *reinterpret_cast<const uint4*>(p);
  const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&u[i]);
    v[2 * i] = __low2float(h);
    v[2 * i + 1] = __high2float(h);
  }
}
template <typename T>
__device__ __forceinline__ void gld8_any(const T* p, float* v, int c, int C, bool vec) {
  if (vec && c + 8 <= C) {
    gld8<T>(p, v);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (c + e < C) ? ldf<T>(p + e) : 0.f;
  }
}

template <typename TX, typename TDY, bool G1>
__global__ void __launch_bounds__(256)
gn_bwd_reduce_kernel(const TX* __restrict__ x, int64_t ldx, const TDY* __restrict__ dy, int64_t lddy,
                     const double* __restrict__ stats, const float* __restrict__ gamma, float* __restrict__ bst,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int HW, int C, int G, float eps,
                     int in_act, const float* __restrict__ beta, int out_act, int vec, int B, int SB) {
  extern __shared__ float sm[];
  float* sdg = sm;          // [C]
  float* sdb = sm + C;      // [C]
  float* sst = sm + 2 * C;  // [2G]
  __shared__ float s_mean[64], s_rstd[64];
  const int cpg = C / G;
  for (int i = threadIdx.x; i < 2 * C + 2 * G; i += blockDim.x) sm[i] = 0.f;
  if (G1 && SB > 1) {
    // One group and all channel octets covered by one pass of the block (C <= 2048): the block walks SB samples and
    // keeps the per-channel sums (dgamma, dbeta) in registers across them, so the shared / global atomics that end a
    // block are paid once per SB samples.  With one sample per block a CelebA-sized training step issued one global
    // atomic per 32 bytes of input (1 M atomics for 33 MB: 36 us at 0.9 TB/s, ncu round 2).
    __syncthreads();
    const int co = (C + 7) >> 3;
    const int tq = min(co, (int)blockDim.x);
    const int lanes = blockDim.x / tq;
    const int pl = threadIdx.x / tq;
    const int q = threadIdx.x - pl * tq;
    const int c = q * 8;
    const bool active = pl < lanes;
    float gm[8], dg[8], db[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      gm[e] = __ldg(gamma + min(c + e, C - 1));
      dg[e] = db[e] = 0.f;
    }
    const int pstride = lanes * gridDim.x;
    for (int sb = 0; sb < SB; ++sb) {
      const int b = blockIdx.y * SB + sb;
      if (b >= B) break;   // block-uniform
      const double inv_n = 1.0 / ((double)HW * C);
      const double s1 = stats[2 * (int64_t)b], s2 = stats[2 * (int64_t)b + 1];
      const double mean_d = s1 * inv_n;
      const float mean = (float)mean_d;
      const float rstd = (float)(1.0 / sqrt(fmax(s2 * inv_n - mean_d * mean_d, 0.0) + (double)eps));
      float a1 = 0.f, a2 = 0.f;
      auto body = [&](const float* xv, const float* d) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float v = xv[e];
          if (in_act) v = act_fwd(v, in_act);
          const float xh = (v - mean) * rstd;
          float de = d[e];
          if (out_act) de *= act_grad(fmaf(xh, gm[e], __ldg(beta + min(c + e, C - 1))), out_act);
          const float t = de * gm[e];
          a1 += t;
          a2 = fmaf(t, xh, a2);
          dg[e] = fmaf(de, xh, dg[e]);
          db[e] += de;
        }
      };
      if (active) {
        int p = blockIdx.x * lanes + pl;
        for (; p + pstride < HW; p += 2 * pstride) {  // two pixels (4 independent 16-byte loads) in flight
          const int64_t pix0 = (int64_t)b * HW + p, pix1 = pix0 + pstride;
          float x0[8], d0[8], x1[8], d1[8];
          gld8_any<TX>(x + pix0 * ldx + c, x0, c, C, vec);
          gld8_any<TDY>(dy + pix0 * lddy + c, d0, c, C, vec);
          gld8_any<TX>(x + pix1 * ldx + c, x1, c, C, vec);
          gld8_any<TDY>(dy + pix1 * lddy + c, d1, c, C, vec);
          body(x0, d0);
          body(x1, d1);
        }
        for (; p < HW; p += pstride) {
          const int64_t pix = (int64_t)b * HW + p;
          float x0[8], d0[8];
          gld8_any<TX>(x + pix * ldx + c, x0, c, C, vec);
          gld8_any<TDY>(dy + pix * lddy + c, d0, c, C, vec);
          body(x0, d0);
        }
      }
      a1 = warp_sum(a1);
      a2 = warp_sum(a2);
      if ((threadIdx.x & 31) == 0 && (a1 != 0.f || a2 != 0.f)) {
        atomicAdd(bst + 2 * (int64_t)b, a1);
        atomicAdd(bst + 2 * (int64_t)b + 1, a2);
      }
    }
    if (active) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (c + e < C) {
          atomicAdd(&sdg[c + e], dg[e]);
          atomicAdd(&sdb[c + e], db[e]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      atomicAdd(dgamma + i, sdg[i]);
      atomicAdd(dbeta + i, sdb[i]);
    }
    return;
  }
  const int b = blockIdx.y;
  if (threadIdx.x < G) {
    const double inv_n = 1.0 / ((double)HW * cpg);
    const double s1 = stats[2 * ((int64_t)b * G + threadIdx.x)], s2 = stats[2 * ((int64_t)b * G + threadIdx.x) + 1];
    const double mean = s1 * inv_n;
    const double var = fmax(s2 * inv_n - mean * mean, 0.0);
    s_mean[threadIdx.x] = (float)mean;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int co = (C + 7) >> 3;
  const int tq = min(co, (int)blockDim.x);
  const int lanes = blockDim.x / tq;
  const int pl = threadIdx.x / tq;
  float a1 = 0.f, a2 = 0.f;  // group sums when there is one group (G1)
  if (pl < lanes) {
    const int pstride = lanes * gridDim.x;
    for (int q = threadIdx.x - pl * tq; q < co; q += tq) {
      const int c = q * 8;
      float gm[8], dg[8], db[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        gm[e] = __ldg(gamma + min(c + e, C - 1));
        dg[e] = db[e] = 0.f;
      }
      auto body = [&](const float* xv, const float* d) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int g = G1 ? 0 : min(c + e, C - 1) / cpg;
          float v = xv[e];
          if (in_act) v = act_fwd(v, in_act);
          const float xh = (v - s_mean[g]) * s_rstd[g];
          float de = d[e];
          if (out_act) de *= act_grad(fmaf(xh, gm[e], __ldg(beta + min(c + e, C - 1))), out_act);
          const float t = de * gm[e];
          if (G1) {
            a1 += t;
            a2 = fmaf(t, xh, a2);
          } else if (c + e < C) {
            atomicAdd(&sst[2 * g], t);
            atomicAdd(&sst[2 * g + 1], t * xh);
          }
          dg[e] = fmaf(de, xh, dg[e]);
          db[e] += de;
        }
      };
      int p = blockIdx.x * lanes + pl;
      for (; p + pstride < HW; p += 2 * pstride) {  // two pixels (4 independent 16-byte loads) in flight
        const int64_t pix0 = (int64_t)b * HW + p, pix1 = pix0 + pstride;
        float x0[8], d0[8], x1[8], d1[8];
        gld8_any<TX>(x + pix0 * ldx + c, x0, c, C, vec);
        gld8_any<TDY>(dy + pix0 * lddy + c, d0, c, C, vec);
        gld8_any<TX>(x + pix1 * ldx + c, x1, c, C, vec);
        gld8_any<TDY>(dy + pix1 * lddy + c, d1, c, C, vec);
        body(x0, d0);
        body(x1, d1);
      }
      for (; p < HW; p += pstride) {
        const int64_t pix = (int64_t)b * HW + p;
        float x0[8], d0[8];
        gld8_any<TX>(x + pix * ldx + c, x0, c, C, vec);
        gld8_any<TDY>(dy + pix * lddy + c, d0, c, C, vec);
        body(x0, d0);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (c + e < C) {
          atomicAdd(&sdg[c + e], dg[e]);
          atomicAdd(&sdb[c + e], db[e]);
        }
    }
  }
  if (G1) {  // padded channels contribute 0 (their dy / x loads are zero-filled)
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&sst[0], a1);
      atomicAdd(&sst[1], a2);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    atomicAdd(dgamma + i, sdg[i]);
    atomicAdd(dbeta + i, sdb[i]);
  }
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(bst + ((int64_t)b * G) * 2 + i, sst[i]);
}

// apply: dx = rstd*(dy*gamma - S1/n - xh*S2/n) [* act'(pre)]  (+ addend)
template <typename TX, typename TDY>
__global__ void __launch_bounds__(256)
gn_bwd_apply_kernel(const TX* __restrict__ x, int64_t ldx, const TDY* __restrict__ dy, int64_t lddy,
                    const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ bst,
                    const float* __restrict__ addend, int64_t ldadd, float* __restrict__ out_f32, int64_t ldo_f32,
                    __nv_bfloat16* __restrict__ out_bf16, int64_t ldo_bf16, int HW, int C, int G, float eps,
                    int in_act, const float* __restrict__ beta, int out_act, int vec) {
  __shared__ float s_mean[64], s_rstd[64], s_c1[64], s_c2[64];
  const int b = blockIdx.y;
  const int cpg = C / G;
  if (threadIdx.x < G) {
    const double inv_n = 1.0 / ((double)HW * cpg);
    const double s1 = stats[2 * ((int64_t)b * G + threadIdx.x)], s2 = stats[2 * ((int64_t)b * G + threadIdx.x) + 1];
    const double mean = s1 * inv_n;
    const double var = fmax(s2 * inv_n - mean * mean, 0.0);
    s_mean[threadIdx.x] = (float)mean;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    s_c1[threadIdx.x] = bst[2 * ((int64_t)b * G + threadIdx.x)] * (float)inv_n;
    s_c2[threadIdx.x] = bst[2 * ((int64_t)b * G + threadIdx.x) + 1] * (float)inv_n;
  }
  __syncthreads();
  const int co = (C + 7) >> 3;
  const int tq = min(co, (int)blockDim.x);
  const int lanes = blockDim.x / tq;
  const int pl = threadIdx.x / tq;
  if (pl >= lanes) return;
  const int pstride = lanes * gridDim.x;
  for (int q = threadIdx.x - pl * tq; q < co; q += tq) {
    const int c = q * 8;
    const bool full = vec && (c + 8 <= C);
    float gm[8], bt[8], mu[8], rs[8], c1[8], c2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ce = min(c + e, C - 1);
      const int g = (G == 1) ? 0 : ce / cpg;
      gm[e] = __ldg(gamma + ce);
      bt[e] = out_act ? __ldg(beta + ce) : 0.f;
      mu[e] = s_mean[g];
      rs[e] = s_rstd[g];
      c1[e] = s_c1[g];
      c2[e] = s_c2[g];
    }
    for (int p = blockIdx.x * lanes + pl; p < HW; p += pstride) {
      const int64_t pix = (int64_t)b * HW + p;
      float pre[8], d[8], dx[8];
      gld8_any<TX>(x + pix * ldx + c, pre, c, C, vec);
      gld8_any<TDY>(dy + pix * lddy + c, d, c, C, vec);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xv = in_act ? act_fwd(pre[e], in_act) : pre[e];
        const float xh = (xv - mu[e]) * rs[e];
        float de = d[e];
        if (out_act) de *= act_grad(fmaf(xh, gm[e], bt[e]), out_act);
        float r = rs[e] * (de * gm[e] - c1[e] - xh * c2[e]);
        if (in_act) r *= act_grad(pre[e], in_act);
        dx[e] = r;
      }
      if (addend) {
        float ad[8];
        gld8_any<float>(addend + pix * ldadd + c, ad, c, C, vec);
#pragma unroll
        for (int e = 0; e < 8; ++e) dx[e] += ad[e];
      }
      if (out_f32) {
        float* op = out_f32 + pix * ldo_f32 + c;
        if (full) {
          *reinterpret_cast<float4*>(op) = make_float4(dx[0], dx[1], dx[2], dx[3]);
          *reinterpret_cast<float4*>(op + 4) = make_float4(dx[4], dx[5], dx[6], dx[7]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (c + e < C) op[e] = dx[e];
        }
      }
      if (out_bf16) {
        __nv_bfloat16* op = out_bf16 + pix * ldo_bf16 + c;
        if (full) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 t = __floats2bfloat162_rn(dx[2 * e], dx[2 * e + 1]);
            w[e] = *reinterpret_cast<uint32_t*>(&t);
          }
          *reinterpret_cast<uint4*>(op) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (c + e < C) op[e] = __float2bfloat16_rn(dx[e]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------ depthwise 7x7 backward (weights)
// block = one sample x 32 channels.  dw[c][tap] += sum_p dy[p] x[p+tap]; db[c] += sum_p dy[p]; dcond[b][c] = sum_p dy[p]
template <int W>
__global__ void __launch_bounds__(256)
dwconv7_wgrad_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ dy, int64_t lddy,
                     float* __restrict__ dw, float* __restrict__ db, float* __restrict__ dcond, int64_t ldc, int C,
                     int H) {
  extern __shared__ float sm[];
  const int HW = H * W;
  float* sx = sm;                       // [HW][33]
  float* sd = sm + (size_t)HW * 33;     // [HW][33]
  float* sred = sd + (size_t)HW * 33;   // [50][32]  (49 taps + bias)
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * 32;
  const int tid = threadIdx.x, cl = tid & 31;
  const int c = c0 + cl;
  const bool c_ok = c < C;
  const int nrow = blockDim.x >> 5;
  for (int p = tid >> 5; p < HW; p += nrow) {
    sx[p * 33 + cl] = c_ok ? __ldg(x + ((int64_t)b * HW + p) * ldx + c) : 0.f;
    sd[p * 33 + cl] = c_ok ? __ldg(dy + ((int64_t)b * HW + p) * lddy + c) : 0.f;
  }
  for (int i = tid; i < 50 * 32; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  float acc[49];
#pragma unroll
  for (int i = 0; i < 49; ++i) acc[i] = 0.f;
  float bsum = 0.f;
  for (int oh = tid >> 5; oh < H; oh += nrow) {
    float dr[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
      dr[i] = sd[(oh * W + i) * 33 + cl];
      bsum += dr[i];
    }
#pragma unroll
    for (int kh = 0; kh < 7; ++kh) {
      const int ih = oh + kh - 3;
      if (ih < 0 || ih >= H) continue;
      float xr[W];
#pragma unroll
      for (int i = 0; i < W; ++i) xr[i] = sx[(ih * W + i) * 33 + cl];
#pragma unroll
      for (int kw = 0; kw < 7; ++kw) {
        float a = 0.f;
#pragma unroll
        for (int ow = 0; ow < W; ++ow) {
          const int iw = ow + kw - 3;
          if (iw >= 0 && iw < W) a = fmaf(dr[ow], xr[iw], a);
        }
        acc[kh * 7 + kw] += a;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 49; ++i) atomicAdd(&sred[i * 32 + cl], acc[i]);
  atomicAdd(&sred[49 * 32 + cl], bsum);
  __syncthreads();
  for (int i = tid; i < 49 * 32; i += blockDim.x) {
    const int tap = i >> 5, ch = i & 31;
    if (c0 + ch < C) atomicAdd(dw + (int64_t)(c0 + ch) * 49 + tap, sred[i]);
  }
  if (tid < 32 && c_ok) {
    const float s = sred[49 * 32 + tid];
    if (db) atomicAdd(db + c, s);
    if (dcond) dcond[(int64_t)b * ldc + c] = s;
  }
}

// Small maps (1x1, 2x2, 4x4: the lower levels of the nets).  The kernel above runs one block per (32-channel chunk,
// sample) and ends every block with 49 x 32 global atomics: 13 M atomics for the 1x1 level of the CelebA net at batch
// 256 (79 us per launch for 0.5 MFLOP).  Here a thread owns a channel and walks a slice of the BATCH with the whole
// map of x and dy in registers (lanes = consecutive channels: coalesced rows), so the atomics are 49 per thread and
// grid slice.  dw must be zero-initialised by the caller (as for the kernel above).
template <int W>
__global__ void __launch_bounds__(128)
dwconv7_wgrad_small_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ dy, int64_t lddy,
                           float* __restrict__ dw, float* __restrict__ db, float* __restrict__ dcond, int64_t ldc,
                           int B, int C) {
  constexpr int HW = W * W;
  constexpr int R = W < 4 ? W - 1 : 3;      // taps dh, dw in [-R, R] see a pixel pair
  constexpr int NT = 2 * R + 1;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc[NT * NT];
#pragma unroll
  for (int i = 0; i < NT * NT; ++i) acc[i] = 0.f;
  float bsum = 0.f;
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    float xv[HW], dv[HW];
#pragma unroll
    for (int pix = 0; pix < HW; ++pix) {
      xv[pix] = __ldg(x + ((int64_t)b * HW + pix) * ldx + c);
      dv[pix] = __ldg(dy + ((int64_t)b * HW + pix) * lddy + c);
    }
    float ds = 0.f;
#pragma unroll
    for (int pix = 0; pix < HW; ++pix) ds += dv[pix];
    bsum += ds;
    if (dcond) dcond[(int64_t)b * ldc + c] = ds;
#pragma unroll
    for (int oh = 0; oh < W; ++oh)
#pragma unroll
      for (int ow = 0; ow < W; ++ow)
#pragma unroll
        for (int ih = 0; ih < W; ++ih)
#pragma unroll
          for (int iw = 0; iw < W; ++iw) {
            const int dh = ih - oh, dwi = iw - ow;
            if (dh >= -R && dh <= R && dwi >= -R && dwi <= R)
              acc[(dh + R) * NT + (dwi + R)] = fmaf(dv[oh * W + ow], xv[ih * W + iw], acc[(dh + R) * NT + (dwi + R)]);
          }
  }
#pragma unroll
  for (int a = 0; a < NT; ++a)
#pragma unroll
    for (int bb = 0; bb < NT; ++bb)
      atomicAdd(dw + (int64_t)c * 49 + (a - R + 3) * 7 + (bb - R + 3), acc[a * NT + bb]);
  if (db) atomicAdd(db + c, bsum);
}

// ------------------------------------------------------------------------------ linear attention backward
// forward (unet_model.py:162-177): qs = softmax_d(q)*scale, ks = softmax_n(k), ctx[d][e] = sum_n ks v, out[n][e] = sum_d ctx qs
__global__ void __launch_bounds__(256)
linear_attn_bwd_kernel(const float* __restrict__ qkv, int64_t ldq, const float* __restrict__ dout, int64_t ldd,
                       __nv_bfloat16* __restrict__ dqkv, int64_t ldg, int n, int heads, float scale) {
  extern __shared__ float sm[];
  float* sq = sm;                       // [n][33]  qs (softmax * scale)
  float* sk = sq + (size_t)n * 33;      // ks
  float* sv = sk + (size_t)n * 33;
  float* sdo = sv + (size_t)n * 33;     // dout, later dks
  float* ctx = sdo + (size_t)n * 33;    // [32][33]
  float* dctx = ctx + 32 * 33;          // [32][33]
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int hid = heads * 32;
  for (int p = warp; p < n; p += nwarp) {
    const float* row = qkv + ((int64_t)b * n + p) * ldq + h * 32 + lane;
    const float qv = row[0];
    float m = qv;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float e = expf(qv - m);
    const float s = warp_sum(e);
    sq[p * 33 + lane] = e / s * scale;
    sk[p * 33 + lane] = row[hid];
    sv[p * 33 + lane] = row[2 * hid];
    sdo[p * 33 + lane] = dout[((int64_t)b * n + p) * ldd + h * 32 + lane];
  }
  __syncthreads();
  for (int d = warp; d < 32; d += nwarp) {
    float m = -INFINITY;
    for (int p = lane; p < n; p += 32) m = fmaxf(m, sk[p * 33 + d]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int p = lane; p < n; p += 32) {
      const float e = expf(sk[p * 33 + d] - m);
      sk[p * 33 + d] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.f / s;
    for (int p = lane; p < n; p += 32) sk[p * 33 + d] *= inv;
  }
  __syncthreads();
  for (int i = tid; i < 1024; i += blockDim.x) {
    const int d = i >> 5, e = i & 31;
    float a = 0.f, g = 0.f;
    for (int p = 0; p < n; ++p) {
      a = fmaf(sk[p * 33 + d], sv[p * 33 + e], a);
      g = fmaf(sdo[p * 33 + e], sq[p * 33 + d], g);
    }
    ctx[d * 33 + e] = a;
    dctx[d * 33 + e] = g;
  }
  __syncthreads();
  for (int p = warp; p < n; p += nwarp) {
    float dqs = 0.f, dv = 0.f, dks = 0.f;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float dok = sdo[p * 33 + k];
      dqs = fmaf(ctx[lane * 33 + k], dok, dqs);              // lane = d, k = e
      dv = fmaf(sk[p * 33 + k], dctx[k * 33 + lane], dv);    // lane = e, k = d
      dks = fmaf(dctx[lane * 33 + k], sv[p * 33 + k], dks);  // lane = d, k = e
    }
    // softmax over d (the lanes) backward: qsm = qs/scale, upstream g = dqs*scale
    const float qsm = sq[p * 33 + lane] / scale;
    const float g = dqs * scale;
    const float dot = warp_sum(qsm * g);
    const float dq = qsm * (g - dot);
    __nv_bfloat16* orow = dqkv + ((int64_t)b * n + p) * ldg + h * 32 + lane;
    orow[0] = __float2bfloat16_rn(dq);
    orow[2 * hid] = __float2bfloat16_rn(dv);
    __syncwarp();
    sdo[p * 33 + lane] = dks;
  }
  __syncthreads();
  for (int d = warp; d < 32; d += nwarp) {
    float dot = 0.f;
    for (int p = lane; p < n; p += 32) dot += sk[p * 33 + d] * sdo[p * 33 + d];
    dot = warp_sum(dot);
    for (int p = lane; p < n; p += 32) {
      const float dk = sk[p * 33 + d] * (sdo[p * 33 + d] - dot);
      dqkv[((int64_t)b * n + p) * ldg + hid + h * 32 + d] = __float2bfloat16_rn(dk);
    }
  }
}

// ------------------------------------------------------------------------------ softmax attention backward
// P = softmax_j(scale * q_i.k_j); out_i = sum_j P_ij v_j.   One block per (head, sample); d walked in chunks of 32.
__global__ void __launch_bounds__(256)
softmax_attn_bwd_kernel(const float* __restrict__ qkv, int64_t ldq, const float* __restrict__ dout, int64_t ldd,
                        __nv_bfloat16* __restrict__ dqkv, int64_t ldg, int n, int dh, int q_off, int k_off,
                        int v_off, int head_stride, float scale) {
  extern __shared__ float sm[];
  float* sP = sm;                           // [n][n+1]
  float* sD = sP + (size_t)n * (n + 1);     // [n][n+1]  dP then dS
  float* sA = sD + (size_t)n * (n + 1);     // [n][33]
  float* sB = sA + (size_t)n * 33;          // [n][33]
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const float* base = qkv + (int64_t)b * n * ldq + h * head_stride;
  const float* dob = dout + (int64_t)b * n * ldd + h * dh;
  __nv_bfloat16* gb = dqkv + (int64_t)b * n * ldg + h * head_stride;
  for (int i = tid; i < n * (n + 1); i += blockDim.x) { sP[i] = 0.f; sD[i] = 0.f; }
  // scores and dP = dout . v^T
  for (int d0 = 0; d0 < dh; d0 += 32) {
    const bool ok = d0 + lane < dh;
    __syncthreads();
    for (int p = warp; p < n; p += nwarp) {
      sA[p * 33 + lane] = ok ? base[(int64_t)p * ldq + q_off + d0 + lane] : 0.f;
      sB[p * 33 + lane] = ok ? base[(int64_t)p * ldq + k_off + d0 + lane] : 0.f;
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += blockDim.x) {
      const int i = idx / n, j = idx - i * n;
      float a = 0.f;
#pragma unroll 8
      for (int d = 0; d < 32; ++d) a = fmaf(sA[i * 33 + d], sB[j * 33 + d], a);
      sP[i * (n + 1) + j] += a;
    }
    __syncthreads();
    for (int p = warp; p < n; p += nwarp) {
      sA[p * 33 + lane] = ok ? dob[(int64_t)p * ldd + d0 + lane] : 0.f;
      sB[p * 33 + lane] = ok ? base[(int64_t)p * ldq + v_off + d0 + lane] : 0.f;
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += blockDim.x) {
      const int i = idx / n, j = idx - i * n;
      float a = 0.f;
#pragma unroll 8
      for (int d = 0; d < 32; ++d) a = fmaf(sA[i * 33 + d], sB[j * 33 + d], a);
      sD[i * (n + 1) + j] += a;
    }
  }
  __syncthreads();
  // softmax rows, then dS = P o (dP - rowsum(P o dP))
  for (int i = warp; i < n; i += nwarp) {
    float m = -INFINITY;
    for (int j = lane; j < n; j += 32) m = fmaxf(m, sP[i * (n + 1) + j] * scale);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int j = lane; j < n; j += 32) {
      const float e = expf(sP[i * (n + 1) + j] * scale - m);
      sP[i * (n + 1) + j] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.f / s;
    float dot = 0.f;
    for (int j = lane; j < n; j += 32) {
      const float pj = sP[i * (n + 1) + j] * inv;
      sP[i * (n + 1) + j] = pj;
      dot += pj * sD[i * (n + 1) + j];
    }
    dot = warp_sum(dot);
    for (int j = lane; j < n; j += 32) sD[i * (n + 1) + j] = sP[i * (n + 1) + j] * (sD[i * (n + 1) + j] - dot) * scale;
  }
  // dv_j = sum_i P_ij dout_i ; dq_i = sum_j dS_ij k_j ; dk_j = sum_i dS_ij q_i
  for (int d0 = 0; d0 < dh; d0 += 32) {
    const bool ok = d0 + lane < dh;
    __syncthreads();
    for (int p = warp; p < n; p += nwarp) sA[p * 33 + lane] = ok ? dob[(int64_t)p * ldd + d0 + lane] : 0.f;
    __syncthreads();
    for (int j = warp; j < n; j += nwarp) {
      float a = 0.f;
      for (int i = 0; i < n; ++i) a = fmaf(sP[i * (n + 1) + j], sA[i * 33 + lane], a);
      if (ok) gb[(int64_t)j * ldg + v_off + d0 + lane] = __float2bfloat16_rn(a);
    }
    __syncthreads();
    for (int p = warp; p < n; p += nwarp) {
      sA[p * 33 + lane] = ok ? base[(int64_t)p * ldq + k_off + d0 + lane] : 0.f;
      sB[p * 33 + lane] = ok ? base[(int64_t)p * ldq + q_off + d0 + lane] : 0.f;
    }
    __syncthreads();
    for (int r = warp; r < n; r += nwarp) {
      float aq = 0.f, ak = 0.f;
      for (int t = 0; t < n; ++t) {
        aq = fmaf(sD[r * (n + 1) + t], sA[t * 33 + lane], aq);  // dq_r = sum_t dS[r][t] k_t
        ak = fmaf(sD[t * (n + 1) + r], sB[t * 33 + lane], ak);  // dk_r = sum_t dS[t][r] q_t
      }
      if (ok) {
        gb[(int64_t)r * ldg + q_off + d0 + lane] = __float2bfloat16_rn(aq);
        gb[(int64_t)r * ldg + k_off + d0 + lane] = __float2bfloat16_rn(ak);
      }
    }
  }
}

// ------------------------------------------------------------------------------ small elementwise helpers
// out = dy * act'(pre)
template <typename TP>
__global__ void act_bwd_kernel(const float* __restrict__ dy, int64_t lddy, const TP* __restrict__ pre, int64_t ldp,
                               float* __restrict__ out_f32, int64_t ldo, __nv_bfloat16* __restrict__ out_bf16,
                               int64_t ldb, int64_t rows, int C, int act) {
  const int64_t total = rows * C;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int64_t r = idx / C;
    const float v = dy[r * lddy + c] * act_grad(ldf<TP>(pre + r * ldp + c), act);
    if (out_f32) out_f32[r * ldo + c] = v;
    if (out_bf16) out_bf16[r * ldb + c] = __float2bfloat16_rn(v);
  }
}
// fp32 NCHW -> bf16 (and fp32) channels-last
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out_bf16, int64_t ldb,
                                    float* __restrict__ out_f32, int64_t ldf_, int B, int C, int HW) {
  const int64_t total = (int64_t)B * HW * C;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int64_t pix = idx / C;
    const int p = (int)(pix % HW);
    const int b = (int)(pix / HW);
    const float v = x[((int64_t)b * C + c) * HW + p];
    if (out_bf16) out_bf16[pix * ldb + c] = __float2bfloat16_rn(v);
    if (out_f32) out_f32[pix * ldf_ + c] = v;
  }
}
// strided fp32: out = a + b (out may alias a); optional bf16 copy of the sum
__global__ void add_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb_,
                           float* __restrict__ out, int64_t ldo, __nv_bfloat16* __restrict__ out_bf16, int64_t ldh,
                           int64_t rows, int C) {
  const int64_t total = rows * C;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int64_t r = idx / C;
    const float v = a[r * lda + c] + (b ? b[r * ldb_ + c] : 0.f);
    if (out) out[r * ldo + c] = v;
    if (out_bf16) out_bf16[r * ldh + c] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------ multi-tensor Adam
// torch.optim.Adam (no amsgrad / weight decay / maximize): exp_avg, exp_avg_sq updates, bias-corrected step.
__global__ void __launch_bounds__(256)
adam_kernel(const sbm_adam_tensor* __restrict__ tensors, const int2* __restrict__ chunks, int chunk_elems, float lr,
            float beta1, float beta2, float eps, float bc1, float bc2_sqrt, float grad_scale,
            const int* __restrict__ step_dev) {
  if (step_dev != nullptr) {  // CUDA-graph replay: the step count lives in device memory (sbm_train_tick advances it)
    const double st = (double)(*step_dev + 1);
    bc1 = (float)(1.0 - pow((double)beta1, st));
    bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, st));
  }
  const int2 ck = chunks[blockIdx.x];
  const sbm_adam_tensor t = tensors[ck.x];
  const int64_t start = (int64_t)ck.y * chunk_elems;
  const int64_t end = min(start + (int64_t)chunk_elems, t.n);
  const float step = lr / bc1;
  auto upd = [&](float g, float& m, float& v, float& p) {
    g *= grad_scale;
    m = beta1 * m + (1.f - beta1) * g;
    v = beta2 * v + (1.f - beta2) * g * g;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    p -= step * (m / denom);
  };
  int64_t i0 = start;
  // 16-byte path: 4 elements per thread, the four streams' loads issued together (28 B/element of HBM traffic)
  if (((reinterpret_cast<uintptr_t>(t.param) | reinterpret_cast<uintptr_t>(t.grad) |
        reinterpret_cast<uintptr_t>(t.exp_avg) | reinterpret_cast<uintptr_t>(t.exp_avg_sq)) & 15) == 0 &&
      (start & 3) == 0) {
    const int64_t nq = (end - start) >> 2;
    float4* p4 = reinterpret_cast<float4*>(t.param + start);
    const float4* g4 = reinterpret_cast<const float4*>(t.grad + start);
    float4* m4 = reinterpret_cast<float4*>(t.exp_avg + start);
    float4* v4 = reinterpret_cast<float4*>(t.exp_avg_sq + start);
    for (int64_t q = threadIdx.x; q < nq; q += blockDim.x) {
      const float4 g = __ldcs(g4 + q);
      float4 m = m4[q], v = v4[q], p = p4[q];
      upd(g.x, m.x, v.x, p.x);
      upd(g.y, m.y, v.y, p.y);
      upd(g.z, m.z, v.z, p.z);
      upd(g.w, m.w, v.w, p.w);
      m4[q] = m;
      v4[q] = v;
      p4[q] = p;
    }
    i0 = start + (nq << 2);
  }
  for (int64_t i = i0 + threadIdx.x; i < end; i += blockDim.x) {
    float m = t.exp_avg[i], v = t.exp_avg_sq[i], p = t.param[i];
    upd(t.grad[i], m, v, p);
    t.exp_avg[i] = m;
    t.exp_avg_sq[i] = v;
    t.param[i] = p;
  }
}

// out[b][c] = sum over the HW pixels of sample b of x[b][p][c]  (gradient of a per-sample row bias, unet_openai.py:303)
template <typename T>
__global__ void __launch_bounds__(256)
colsum_per_sample_kernel(const T* __restrict__ x, int64_t ld, int HW, int C, float* __restrict__ out, int64_t ldo) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  __shared__ float red[8][33];
  float acc = 0.f;
  if (c < C)
    for (int p = threadIdx.x >> 5; p < HW; p += 8) acc += ldf<T>(x + ((int64_t)b * HW + p) * ld + c);
  red[threadIdx.x >> 5][threadIdx.x & 31] = acc;
  __syncthreads();
  if (threadIdx.x < 32 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    out[(int64_t)b * ldo + c] = t;
  }
}

// backward of nearest-neighbour 2x upsampling (unet_openai.py:185): out[b,i,j,c] = sum of the 2x2 block of dy
__global__ void __launch_bounds__(256)
upsample2x_bwd_kernel(const float* __restrict__ dy, int64_t lddy, float* __restrict__ out, int64_t ldo,
                      __nv_bfloat16* __restrict__ out_bf16, int64_t ldb, int B, int H, int W, int C) {
  const int64_t total = (int64_t)B * H * W * C;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int64_t pix = idx / C;
    const int j = (int)(pix % W);
    const int i = (int)((pix / W) % H);
    const int64_t b = pix / ((int64_t)W * H);
    const float* r0 = dy + ((b * 2 * H + 2 * i) * 2 * W + 2 * j) * lddy + c;
    const float* r1 = r0 + (int64_t)2 * W * lddy;
    const float v = r0[0] + r0[lddy] + r1[0] + r1[lddy];
    if (out) out[pix * ldo + c] = v;
    if (out_bf16) out_bf16[pix * ldb + c] = __float2bfloat16_rn(v);
  }
}

// utils.py:79-90 update_ema: ema = ema*decay + param*(1 - decay), all tensors in one launch
__global__ void __launch_bounds__(256)
ema_kernel(const sbm_ema_tensor* __restrict__ tensors, const int2* __restrict__ chunks, int chunk_elems, float decay) {
  const int2 ck = chunks[blockIdx.x];
  const sbm_ema_tensor t = tensors[ck.x];
  const int64_t start = (int64_t)ck.y * chunk_elems;
  const int64_t end = min(start + (int64_t)chunk_elems, t.n);
  const float a = 1.f - decay;
  for (int64_t i = start + threadIdx.x; i < end; i += blockDim.x) t.ema[i] = t.ema[i] * decay + a * t.src[i];
}

static int egrid(int64_t n) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8));
}

// ------------------------------------------------------------------------------ scale-shift modulation, backward
// y = act(u), u = n * (1 + scale[b][c]) + shift[b][c]  (ResBlock(use_scale_shift_norm=True), unet_openai.py:296-300; n
// = GroupNorm32(h)).  Given dy:  du = dy * act'(u);  dn = du * (1 + scale);  dscale[b][c] += sum_pix du * n;
// dshift[b][c] += sum_pix du.  Block = (pixel chunk, sample); thread = channel (consecutive threads -> consecutive
// channels), sums over the chunk's pixels in registers, one global atomic pair per (thread, channel).
__global__ void __launch_bounds__(256)
scale_shift_bwd_kernel(const float* __restrict__ n, int64_t ldn, const float* __restrict__ dy, int64_t lddy,
                       const float* __restrict__ scale, const float* __restrict__ shift, int64_t ld_mod, int act,
                       float* __restrict__ dn, int64_t lddn, float* __restrict__ dscale, float* __restrict__ dshift,
                       int64_t ld_dmod, int HW, int C, int ppb) {
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * ppb, p1 = min(p0 + ppb, HW);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float s1 = 1.f + __ldg(scale + (int64_t)b * ld_mod + c), t = __ldg(shift + (int64_t)b * ld_mod + c);
    float as = 0.f, at = 0.f;
    for (int p = p0; p < p1; ++p) {
      const int64_t pix = (int64_t)b * HW + p;
      const float nv = n[pix * ldn + c];
      const float du = dy[pix * lddy + c] * act_grad(fmaf(nv, s1, t), act);
      dn[pix * lddn + c] = du * s1;
      as = fmaf(du, nv, as);
      at += du;
    }
    atomicAdd(dscale + (int64_t)b * ld_dmod + c, as);
    atomicAdd(dshift + (int64_t)b * ld_dmod + c, at);
  }
}

}  // namespace sbm

using namespace sbm;

extern "C" {

int sbm_colsum(const void* x, int32_t dtype, int64_t ld, int64_t rows, int32_t C, float* out, void* stream) {
  SBM_CHECK_ARG(x && out && rows > 0 && C > 0, "sbm_colsum: bad args");
  dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((rows + 63) / 64, 512)), (unsigned)std::min((C + 31) / 32, 64));
  if (dtype == SBM_F32) colsum_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, ld, rows, C, out);
  else colsum_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld, rows, C, out);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_b();
  return 0;
}

int sbm_colsum_per_sample(const void* x, int32_t dtype, int64_t ld, int32_t B, int32_t HW, int32_t C, float* out,
                          int64_t ldo, void* stream) {
  SBM_CHECK_ARG(x && out && B > 0 && HW > 0 && C > 0, "sbm_colsum_per_sample: bad args");
  dim3 grid((C + 31) / 32, B);
  if (dtype == SBM_F32)
    colsum_per_sample_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, ld, HW, C, out, ldo);
  else
    colsum_per_sample_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ld, HW, C,
                                                                                    out, ldo);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_b();
  return 0;
}

int sbm_scale_shift_bwd(const float* n, int64_t ldn, const float* dy, int64_t lddy, const float* scale,
                        const float* shift, int64_t ld_mod, int32_t act, float* dn, int64_t lddn, float* dscale,
                        float* dshift, int64_t ld_dmod, int32_t B, int32_t HW, int32_t C, void* stream) {
  SBM_CHECK_ARG(n && dy && scale && shift && dn && dscale && dshift && B > 0 && HW > 0 && C > 0,
                "sbm_scale_shift_bwd: bad args");
  // enough blocks for ~4 per SM, at least 8 pixels each
  const int want = std::max(1, sm_count() * 4 / B);
  const int ppb = std::max(8, (HW + want - 1) / want);
  dim3 grid((HW + ppb - 1) / ppb, B);
  scale_shift_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, ldn, dy, lddy, scale, shift, ld_mod, act, dn, lddn,
                                                                 dscale, dshift, ld_dmod, HW, C, ppb);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_b();
  return 0;
}

int sbm_upsample_nearest2x_bwd(const float* dy, int64_t lddy, float* out, int64_t ldo, void* out_bf16, int64_t ldb,
                               int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  SBM_CHECK_ARG(dy && (out || out_bf16) && B > 0 && H > 0 && W > 0 && C > 0, "sbm_upsample_nearest2x_bwd: bad args");
  upsample2x_bwd_kernel<<<egrid((int64_t)B * H * W * C), 256, 0, (cudaStream_t)stream>>>(
      dy, lddy, out, ldo, (__nv_bfloat16*)out_bf16, ldb, B, H, W, C);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_b();
  return 0;
}

int sbm_groupnorm_bwd(const void* x, int32_t x_dtype, int64_t ldx, const void* dy, int32_t dy_dtype, int64_t lddy,
                      const double* stats, const float* gamma, float* bst, float* dgamma, float* dbeta,
                      const float* addend, int64_t ldadd, float* out_f32, int64_t ldo_f32, void* out_bf16,
                      int64_t ldo_bf16, int32_t B, int32_t HW, int32_t C, int32_t G, float eps, int32_t in_act,
                      const float* beta, int32_t out_act, void* stream) {
  SBM_CHECK_ARG(x && dy && stats && gamma && bst && dgamma && dbeta && (out_f32 || out_bf16), "sbm_groupnorm_bwd: null");
  SBM_CHECK_ARG(out_act == 0 || beta != nullptr, "sbm_groupnorm_bwd: an output activation needs beta");
  SBM_CHECK_ARG(B > 0 && G > 0 && G <= 64 && C % G == 0, "sbm_groupnorm_bwd: bad sizes");
  // block = (channel octets) x (pixel lanes), like the forward apply kernel
  const int co = (C + 7) / 8;
  const int lanes = std::max(1, 256 / std::min(co, 256));
  int chunks = (int)std::min<int64_t>((HW + lanes - 1) / lanes, std::max<int64_t>(1, (int64_t)sm_count() * 8 / B));
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, B);
  // reduce pass, one group, every channel octet owned by one thread of the block: SB samples per block, sized so that
  // the grid still holds about two blocks per SM
  int SB = 1;
  if (G == 1 && co <= 256) SB = (int)std::max<int64_t>(1, std::min<int64_t>(16, (int64_t)chunks * B / (2 * (int64_t)sm_count())));
  dim3 grid_r(chunks, (B + SB - 1) / SB);
  const size_t smem = (size_t)(2 * C + 2 * G) * sizeof(float);
  SBM_CHECK_ARG(smem <= 48 * 1024, "sbm_groupnorm_bwd: C=%d too large", C);
  const int xs = x_dtype == SBM_F32 ? 4 : 2, ds = dy_dtype == SBM_F32 ? 4 : 2;
  auto al16 = [](const void* p, int64_t ld, int esz) {
    return p == nullptr || (((reinterpret_cast<uintptr_t>(p) & 15) == 0) && ((ld * esz) % 16 == 0));
  };
  const int vec = al16(x, ldx, xs) && al16(dy, lddy, ds) && al16(addend, ldadd, 4) && al16(out_f32, ldo_f32, 4) &&
                  al16(out_bf16, ldo_bf16, 2);
  cudaStream_t s = (cudaStream_t)stream;
#define SBM_GNB(TX, TDY)                                                                                            \
  do {                                                                                                              \
    if (G == 1)                                                                                                     \
      gn_bwd_reduce_kernel<TX, TDY, true><<<grid_r, 256, smem, s>>>((const TX*)x, ldx, (const TDY*)dy, lddy, stats,   \
                                                                    gamma, bst, dgamma, dbeta, HW, C, G, eps,        \
                                                                    in_act, beta, out_act, vec, B, SB);             \
    else                                                                                                            \
      gn_bwd_reduce_kernel<TX, TDY, false><<<grid, 256, smem, s>>>((const TX*)x, ldx, (const TDY*)dy, lddy, stats,    \
                                                                   gamma, bst, dgamma, dbeta, HW, C, G, eps, in_act, \
                                                                   beta, out_act, vec, B, 1);                       \
    gn_bwd_apply_kernel<TX, TDY><<<grid, 256, 0, s>>>((const TX*)x, ldx, (const TDY*)dy, lddy, stats, gamma, bst,     \
                                                      addend, ldadd, out_f32, ldo_f32, (__nv_bfloat16*)out_bf16,     \
                                                      ldo_bf16, HW, C, G, eps, in_act, beta, out_act, vec);         \
  } while (0)
  if (x_dtype == SBM_F32 && dy_dtype == SBM_F32) SBM_GNB(float, float);
  else if (x_dtype == SBM_BF16 && dy_dtype == SBM_F32) SBM_GNB(__nv_bfloat16, float);
  else if (x_dtype == SBM_F32 && dy_dtype == SBM_BF16) SBM_GNB(float, __nv_bfloat16);
  else SBM_GNB(__nv_bfloat16, __nv_bfloat16);
#undef SBM_GNB
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_b();
  count_launch_b();
  return 0;
}

int sbm_dwconv7_wgrad(const float* x, int64_t ldx, const float* dy, int64_t lddy, float* dw, float* db, float* dcond,
                      int64_t ldc, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  SBM_CHECK_ARG(x && dy && dw && B > 0 && C > 0, "sbm_dwconv7_wgrad: bad args");
  const size_t smem = ((size_t)2 * H * W * 33 + 50 * 32) * sizeof(float);
  SBM_CHECK_ARG(smem <= 200 * 1024, "sbm_dwconv7_wgrad: map too large");
  dim3 grid((C + 31) / 32, B);
  const int threads = 32 * std::max(1, std::min(8, H));
  cudaStream_t s = (cudaStream_t)stream;
#define SBM_DWW(WW)                                                                                                \
  do {                                                                                                             \
    static size_t conf = 0;                                                                                        \
    if (smem > 48 * 1024 && smem > conf) {                                                                         \
      SBM_CUDA_OK(cudaFuncSetAttribute(dwconv7_wgrad_kernel<WW>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                       (int)smem));                                                                \
      conf = smem;                                                                                                 \
    }                                                                                                              \
    dwconv7_wgrad_kernel<WW><<<grid, threads, smem, s>>>(x, ldx, dy, lddy, dw, db, dcond, ldc, C, H);               \
  } while (0)
  static const bool small_env = [] { const char* e = getenv("SBM_DWWGRAD_SMALL"); return e ? atoi(e) != 0 : true; }();
  if (small_env && H == W && W <= 4) {
    // slices of the batch: enough blocks to fill the SMs, few enough that the final atomics stay cheap
    const int cblocks = (C + 127) / 128;
    const int gy = std::max(1, std::min((int)B, (2 * sm_count() + cblocks - 1) / cblocks));
    dim3 g2(cblocks, gy);
    if (W == 1) dwconv7_wgrad_small_kernel<1><<<g2, 128, 0, s>>>(x, ldx, dy, lddy, dw, db, dcond, ldc, B, C);
    else if (W == 2) dwconv7_wgrad_small_kernel<2><<<g2, 128, 0, s>>>(x, ldx, dy, lddy, dw, db, dcond, ldc, B, C);
    else dwconv7_wgrad_small_kernel<4><<<g2, 128, 0, s>>>(x, ldx, dy, lddy, dw, db, dcond, ldc, B, C);
    SBM_CUDA_OK(cudaGetLastError());
    count_launch_b();
    return 0;
  }
  if (W == 16) SBM_DWW(16);
  else if (W == 8) SBM_DWW(8);
  else if (W == 4) SBM_DWW(4);
  else if (W == 2) SBM_DWW(2);
  else if (W == 1) SBM_DWW(1);
  else SBM_CHECK_ARG(false, "sbm_dwconv7_wgrad: unsupported width %d", W);
#undef SBM_DWW
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_b();
  return 0;
}

int sbm_linear_attn_bwd(const float* qkv, int64_t ldq, const float* dout, int64_t ldd, void* dqkv, int64_t ldg,
                        int32_t B, int32_t n, int32_t heads, float scale, void* stream) {
  SBM_CHECK_ARG(qkv && dout && dqkv && B > 0 && n > 0, "sbm_linear_attn_bwd: bad args");
  const size_t smem = ((size_t)4 * n * 33 + 2 * 32 * 33) * sizeof(float);
  SBM_CHECK_ARG(smem <= 200 * 1024, "sbm_linear_attn_bwd: n=%d too large", n);
  static size_t conf = 0;
  if (smem > 48 * 1024 && smem > conf) {
    SBM_CUDA_OK(cudaFuncSetAttribute(linear_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = smem;
  }
  linear_attn_bwd_kernel<<<dim3(heads, B), 256, smem, (cudaStream_t)stream>>>(qkv, ldq, dout, ldd, (__nv_bfloat16*)dqkv,
                                                                              ldg, n, heads, scale);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_b();
  return 0;
}

int sbm_softmax_attn_bwd(const float* qkv, int64_t ldq, const float* dout, int64_t ldd, void* dqkv, int64_t ldg,
                         int32_t B, int32_t n, int32_t heads, int32_t dh, int32_t q_off, int32_t k_off, int32_t v_off,
                         int32_t head_stride, float scale, void* stream) {
  SBM_CHECK_ARG(qkv && dout && dqkv && B > 0 && n > 0, "sbm_softmax_attn_bwd: bad args");
  const size_t smem = ((size_t)2 * n * (n + 1) + 2 * (size_t)n * 33) * sizeof(float);
  SBM_CHECK_ARG(smem <= 200 * 1024, "sbm_softmax_attn_bwd: n=%d too large", n);
  static size_t conf = 0;
  if (smem > 48 * 1024 && smem > conf) {
    SBM_CUDA_OK(cudaFuncSetAttribute(softmax_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = smem;
  }
  softmax_attn_bwd_kernel<<<dim3(heads, B), 256, smem, (cudaStream_t)stream>>>(
      qkv, ldq, dout, ldd, (__nv_bfloat16*)dqkv, ldg, n, dh, q_off, k_off, v_off, head_stride, scale);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_b();
  return 0;
}

int sbm_act_bwd(const float* dy, int64_t lddy, const void* pre, int32_t pre_dtype, int64_t ldp, float* out_f32,
                int64_t ldo, void* out_bf16, int64_t ldb, int64_t rows, int32_t C, int32_t act, void* stream) {
  SBM_CHECK_ARG(dy && pre && (out_f32 || out_bf16) && rows > 0 && C > 0, "sbm_act_bwd: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  if (pre_dtype == SBM_F32)
    act_bwd_kernel<float><<<egrid(rows * C), 256, 0, s>>>(dy, lddy, (const float*)pre, ldp, out_f32, ldo,
                                                          (__nv_bfloat16*)out_bf16, ldb, rows, C, act);
  else
    act_bwd_kernel<__nv_bfloat16><<<egrid(rows * C), 256, 0, s>>>(dy, lddy, (const __nv_bfloat16*)pre, ldp, out_f32, ldo,
                                                                  (__nv_bfloat16*)out_bf16, ldb, rows, C, act);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_b();
  return 0;
}

int sbm_nchw_to_nhwc(const float* x, void* out_bf16, int64_t ldb, float* out_f32, int64_t ldf, int32_t B, int32_t C,
                     int32_t HW, void* stream) {
  SBM_CHECK_ARG(x && (out_bf16 || out_f32) && B > 0 && C > 0 && HW > 0, "sbm_nchw_to_nhwc: bad args");
  nchw_to_nhwc_kernel<<<egrid((int64_t)B * C * HW), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)out_bf16, ldb,
                                                                                    out_f32, ldf, B, C, HW);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_b();
  return 0;
}

int sbm_add(const float* a, int64_t lda, const float* b, int64_t ldb, float* out, int64_t ldo, void* out_bf16,
            int64_t ldh, int64_t rows, int32_t C, void* stream) {
  SBM_CHECK_ARG(a && (out || out_bf16) && rows > 0 && C > 0, "sbm_add: bad args");
  add_kernel<<<egrid(rows * C), 256, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, out, ldo, (__nv_bfloat16*)out_bf16, ldh,
                                                                rows, C);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_b();
  return 0;
}

int sbm_ema_step(const sbm_ema_tensor* tensors_dev, const int32_t* chunks_dev, int32_t n_chunks, int32_t chunk_elems,
                 float decay, void* stream) {
  SBM_CHECK_ARG(tensors_dev && chunks_dev && n_chunks > 0 && chunk_elems > 0, "sbm_ema_step: bad args");
  ema_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(tensors_dev, (const int2*)chunks_dev, chunk_elems, decay);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_b();
  return 0;
}

int sbm_adam_step(const sbm_adam_tensor* tensors_dev, const int32_t* chunks_dev, int32_t n_chunks,
                  int32_t chunk_elems, float lr, float beta1, float beta2, float eps, int32_t step, float grad_scale,
                  const int32_t* step_dev, void* stream) {
  SBM_CHECK_ARG(tensors_dev && chunks_dev && n_chunks > 0 && chunk_elems > 0 && (step >= 1 || step_dev),
                "sbm_adam_step: bad args");
  if (step < 1) step = 1;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(tensors_dev, (const int2*)chunks_dev, chunk_elems, lr, beta1,
                                                          beta2, eps, (float)bc1, (float)sqrt(bc2), grad_scale,
                                                          (const int*)step_dev);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_b();
  return 0;
}

}  // extern "C"
