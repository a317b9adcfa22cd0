// Memory-bound score-net operators (channels-last, fp32 math, bf16 GEMM operands out):
//   * stem im2col                      (unet_model.py:208,287 / unet_openai.py:441)
//   * depthwise 7x7 + bias + time-embedding condition + GroupNorm statistics (unet_model.py:103,116-121)
//   * GroupNorm apply (+affine, +SiLU, +residual)  (unet_model.py:106,109,160,183; unet_openai.py:10-12,63)
//   * group statistics                 (same GroupNorm sites, when the producer cannot emit them)
//   * sinusoidal time embeddings       (unet_model.py:40-47; unet_openai.py:66-83)
//   * linear attention / softmax attention cores (unet_model.py:135-149,162-177; unet_openai.py:345-358)
// All reductions are fp32 per thread, combined in fp64 (atomicAdd(double)) so the E[x^2]-E[x]^2 form is safe.
#include <atomic>
#include <cstdlib>

#include "../../include/sbmae_b200.h"
#include "common.cuh"

namespace sbm {
extern std::atomic<unsigned long long> g_launches;
static inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// 4 consecutive channels of a qkv row: fp32 (16 bytes) or bf16 (8 bytes; the to_qkv GEMM then writes and this kernel
// reads half the bytes -- at 16x16 the pair is HBM-bound on exactly those)
__device__ __forceinline__ float4 load_qkv4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 load_qkv4(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
}


// ------------------------------------------------------------------------------ stem im2col
// x: fp32 NCHW [B,C,H,W]  ->  a: bf16 [B*H*W, ldk], column k = (c*KH + kh)*KW + kw (matches weight.view(O,-1))
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ a, int B, int C, int H, int W, int KH, int KW,
                   int ldk) {
  // one thread = 8 consecutive im2col columns (one 16-byte store) of one pixel; the 49*C taps of a pixel come from a
  // (KH x KW x C) window that stays L1/L2 resident across the threads of the pixel.
  const int K = C * KH * KW;
  const int groups = ldk >> 3;
  const int HWp = H * W;
  const int64_t total = (int64_t)B * HWp * groups;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int gq = (int)(idx % groups);
    const int64_t pix = idx / groups;
    const int pin = (int)(pix % HWp);
    const int b = (int)(pix / HWp);
    const int h = pin / W, w = pin - h * W;
    const float* xb = x + (int64_t)b * C * HWp;
    uint32_t packed[4];
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      float v[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int k = gq * 8 + e2 * 2 + u;
        float val = 0.f;
        if (k < K) {
          const int kw = k % KW;
          const int r = k / KW;
          const int kh = r % KH;
          const int c = r / KH;
          const int ih = h + kh - KH / 2, iw = w + kw - KW / 2;
          if (ih >= 0 && ih < H && iw >= 0 && iw < W) val = __ldg(xb + (c * H + ih) * W + iw);
        }
        v[u] = val;
      }
      __nv_bfloat162 t = __floats2bfloat162_rn(v[0], v[1]);
      packed[e2] = *reinterpret_cast<uint32_t*>(&t);
    }
    *reinterpret_cast<uint4*>(a + pix * ldk + gq * 8) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  }
}

// ------------------------------------------------------------------------------ depthwise 7x7
// One block = one sample x one chunk of 32 channels x all pixels.  The input slab (H*W x 32 channels, fp32)
// and the 49x32 filter taps are staged in shared memory, so global memory is read exactly once.
// h = dwconv7(x) + bias[c] + cond[b][c];  stats[b] += (sum h, sum h^2)
constexpr int kDwCh = 32;
__global__ void __launch_bounds__(256)
dwconv7_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w, const float* __restrict__ bias,
               const float* __restrict__ cond, int64_t ldc, float* __restrict__ out, int64_t ldo,
               double* __restrict__ stats, int C, int H, int W) {
  extern __shared__ float sm[];
  const int HW = H * W;
  float* sx = sm;                      // [HW][33]
  float* sw = sm + (size_t)HW * 33;    // [49][32]
  __shared__ float red[32];
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * kDwCh;
  const int tid = threadIdx.x;
  const int cl = tid & 31;
  const int c = c0 + cl;
  const bool c_ok = c < C;
  for (int p = tid >> 5; p < HW; p += blockDim.x >> 5)
    sx[p * 33 + cl] = c_ok ? __ldg(x + ((int64_t)b * HW + p) * ldx + c) : 0.f;
  for (int i = tid; i < 49 * kDwCh; i += blockDim.x) {
    const int ch = i / 49, tap = i - ch * 49;  // consecutive threads read consecutive taps of one channel
    sw[tap * kDwCh + ch] = (c0 + ch < C) ? __ldg(w + (int64_t)(c0 + ch) * 49 + tap) : 0.f;
  }
  __syncthreads();
  const float add = c_ok ? (bias ? __ldg(bias + c) : 0.f) + (cond ? __ldg(cond + (int64_t)b * ldc + c) : 0.f) : 0.f;
  float s1 = 0.f, s2 = 0.f;
  for (int p = tid >> 5; p < HW; p += blockDim.x >> 5) {
    const int ph = p / W, pw = p - ph * W;
    float acc = 0.f;
    const int kh0 = max(0, 3 - ph), kh1 = min(7, H + 3 - ph);
    const int kw0 = max(0, 3 - pw), kw1 = min(7, W + 3 - pw);
    for (int kh = kh0; kh < kh1; ++kh) {
      const float* row = sx + ((ph + kh - 3) * W + (pw - 3)) * 33 + cl;
      const float* wr = sw + (kh * 7) * kDwCh + cl;
      for (int kw = kw0; kw < kw1; ++kw) acc = fmaf(row[kw * 33], wr[kw * kDwCh], acc);
    }
    acc += add;
    if (c_ok) {
      out[((int64_t)b * HW + p) * ldo + c] = acc;
      s1 += acc;
      s2 += acc * acc;
    }
  }
  if (stats != nullptr) {
    const float t1 = block_sum(s1, red);
    const float t2 = block_sum(s2, red);
    if (tid == 0) {
      atomicAdd(stats + 2 * (int64_t)b, (double)t1);
      atomicAdd(stats + 2 * (int64_t)b + 1, (double)t2);
    }
  }
}


// 1x1 maps (the bottom level of the CelebA net): only the centre tap sees a pixel, so the "convolution" is
// h[b][c] = x[b][c] * w[c][3][3] + bias[c] + cond[b][c].  One warp per sample; the tiled kernels spent 24 us per
// launch on staging for it (64 blocks, one 32-channel chunk each).
template <typename TOut>
__global__ void __launch_bounds__(256)
dwconv7_point_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                     const float* __restrict__ bias, const float* __restrict__ cond, int64_t ldc,
                     TOut* __restrict__ out, int64_t ldo, double* __restrict__ stats, int B, int C) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < C; c += 32) {
    float h = fmaf(__ldg(x + (int64_t)b * ldx + c), __ldg(w + (int64_t)c * 49 + 24), bias ? __ldg(bias + c) : 0.f);
    if (cond != nullptr) h += __ldg(cond + (int64_t)b * ldc + c);
    if constexpr (sizeof(TOut) == 2) {
      const __nv_bfloat16 r = __float2bfloat16_rn(h);
      out[(int64_t)b * ldo + c] = r;
      h = __bfloat162float(r);   // the statistics describe the tensor the next GEMM reads
    } else {
      out[(int64_t)b * ldo + c] = h;
    }
    s1 += h;
    s2 = fmaf(h, h, s2);
  }
  if (stats != nullptr) {
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      atomicAdd(stats + 2 * (int64_t)b, (double)s1);
      atomicAdd(stats + 2 * (int64_t)b + 1, (double)s2);
    }
  }
}

// Register-tiled variant for W in {1,2,4,8,16}: one thread = one (channel, output row); the 49 taps live in
// registers and every staged input value is read from shared memory once per kernel row (7 LDS per input
// instead of 49), so the kernel is FMA-bound instead of LDS-bound.
template <int W>
__global__ void __launch_bounds__(256)
dwconv7_rows_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                    const float* __restrict__ bias, const float* __restrict__ cond, int64_t ldc,
                    float* __restrict__ out, int64_t ldo, double* __restrict__ stats, int C, int H, int flip,
                    const float* __restrict__ addend, int64_t ldadd) {
  extern __shared__ float sm[];
  const int HW = H * W;
  float* sx = sm;  // [HW][33]
  __shared__ float red[32];
  const int b = blockIdx.y;
  const int c0 = blockIdx.x * kDwCh;
  const int tid = threadIdx.x;
  const int cl = tid & 31;
  const int c = c0 + cl;
  const bool c_ok = c < C;
  const int nrow = blockDim.x >> 5;
  {
    // coalesced slab load with 4 independent loads in flight per thread
    const float* xb = x + (int64_t)b * HW * ldx + c;
    int p = tid >> 5;
    for (; p + 3 * nrow < HW; p += 4 * nrow) {
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
      if (c_ok) {
        v0 = __ldg(xb + (int64_t)p * ldx);
        v1 = __ldg(xb + (int64_t)(p + nrow) * ldx);
        v2 = __ldg(xb + (int64_t)(p + 2 * nrow) * ldx);
        v3 = __ldg(xb + (int64_t)(p + 3 * nrow) * ldx);
      }
      sx[p * 33 + cl] = v0;
      sx[(p + nrow) * 33 + cl] = v1;
      sx[(p + 2 * nrow) * 33 + cl] = v2;
      sx[(p + 3 * nrow) * 33 + cl] = v3;
    }
    for (; p < HW; p += nrow) sx[p * 33 + cl] = c_ok ? __ldg(xb + (int64_t)p * ldx) : 0.f;
  }
  float wr[49];
#pragma unroll
  for (int i = 0; i < 49; ++i) wr[i] = c_ok ? __ldg(w + (int64_t)c * 49 + (flip ? 48 - i : i)) : 0.f;
  const float add = c_ok ? (bias ? __ldg(bias + c) : 0.f) + (cond ? __ldg(cond + (int64_t)b * ldc + c) : 0.f) : 0.f;
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
  for (int oh = tid >> 5; oh < H; oh += nrow) {
    float acc[W];
#pragma unroll
    for (int i = 0; i < W; ++i) acc[i] = add;
#pragma unroll
    for (int kh = 0; kh < 7; ++kh) {
      const int ih = oh + kh - 3;
      if (ih < 0 || ih >= H) continue;
      const float* row = sx + (ih * W) * 33 + cl;
#pragma unroll
      for (int iw = 0; iw < W; ++iw) {
        const float v = row[iw * 33];
#pragma unroll
        for (int kw = 0; kw < 7; ++kw) {
          const int ow = iw - kw + 3;
          if (ow >= 0 && ow < W) acc[ow] = fmaf(v, wr[kh * 7 + kw], acc[ow]);
        }
      }
    }
    if (c_ok) {
      float* op = out + ((int64_t)b * HW + oh * W) * ldo + c;
      if (addend != nullptr) {
        const float* ap = addend + ((int64_t)b * HW + oh * W) * ldadd + c;
#pragma unroll
        for (int i = 0; i < W; ++i) acc[i] += ap[(int64_t)i * ldadd];
      }
#pragma unroll
      for (int i = 0; i < W; ++i) {
        op[(int64_t)i * ldo] = acc[i];
        s1 += acc[i];
        s2 += acc[i] * acc[i];
      }
    }
  }
  if (stats != nullptr) {
    const float t1 = block_sum(s1, red);
    const float t2 = block_sum(s2, red);
    if (tid == 0) {
      atomicAdd(stats + 2 * (int64_t)b, (double)t1);
      atomicAdd(stats + 2 * (int64_t)b + 1, (double)t2);
    }
  }
}

// Pipelined variant (the one the forward pass uses): a block owns ONE 32-channel chunk (its 49 taps stay in registers)
// and walks the samples; the input slab of the next step is fetched with cp.async (all 16-byte requests of the slab in
// flight at once) while the current one is convolved, so the kernel is FMA-bound instead of load-latency-bound.
// Maps smaller than 8 rows pack 8/H samples per step so that all 8 warps stay busy.  Output fp32 or bf16.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int W, typename TOut>
__global__ void __launch_bounds__(256, (W <= 8 ? 3 : 2))
dwconv7_pipe_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                    const float* __restrict__ bias, const float* __restrict__ cond, int64_t ldc,
                    TOut* __restrict__ out, int64_t ldo, double* __restrict__ stats, int B, int C, int H, int spb,
                    int flip, const float* __restrict__ addend, int64_t ldadd) {
  extern __shared__ __align__(16) float sm[];
  const int HW = H * W;
  const int slab = spb * HW * kDwCh;           // floats per buffer (spb samples per step)
  const int c0 = blockIdx.x * kDwCh;
  const int tid = threadIdx.x;
  const int cl = tid & 31, warp = tid >> 5;
  const int c = c0 + cl;
  const bool c_ok = c < C;
  const int nsteps = (B + spb - 1) / spb;
  const int log_hw = 31 - __clz(HW);           // maps are powers of two

  auto prefetch = [&](int step, float* buf) {
    // slab = spb samples x HW pixels x 32 channels: 8 sixteen-byte chunks per pixel, all requests in flight at once
    const int chunks = spb * HW * 8;
    for (int i = tid; i < chunks; i += 256) {
      const int q = i & 7, pix = i >> 3;
      const int b = step * spb + (pix >> log_hw);
      if (b < B && c0 + q * 4 < C)
        cp_async16(buf + pix * kDwCh + q * 4, x + ((int64_t)step * spb * HW + pix) * ldx + c0 + q * 4);
    }
    cp_async_commit();
  };

  float wr[49];
#pragma unroll
  for (int i = 0; i < 49; ++i) wr[i] = c_ok ? __ldg(w + (int64_t)c * 49 + (flip ? 48 - i : i)) : 0.f;
  const float bias_c = (c_ok && bias) ? __ldg(bias + c) : 0.f;

  int step = blockIdx.y;
  if (step < nsteps) prefetch(step, sm);
  int cur = 0;
  for (; step < nsteps; step += gridDim.y) {
    const int nxt = step + gridDim.y;
    if (nxt < nsteps) {
      prefetch(nxt, sm + (cur ^ 1) * slab);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    // a warp owns a contiguous run of output rows, i.e. rows of ONE sample when H >= 8 / whole samples otherwise, so
    // the GroupNorm statistics need one warp reduction per sample instead of one per row
    {
      const int rows_total = spb * H;
      const int per_warp = (rows_total + 7) >> 3;
      const int r_begin = warp * per_warp, r_end = min(rows_total, r_begin + per_warp);
      int cur_ls = -1;
      float s1 = 0.f, s2 = 0.f, add = 0.f;
      auto flush = [&](int ls_) {
        if (stats != nullptr && ls_ >= 0) {
          const float t1 = warp_sum(s1), t2 = warp_sum(s2);
          if (cl == 0) {
            const int bb = step * spb + ls_;
            atomicAdd(stats + 2 * (int64_t)bb, (double)t1);
            atomicAdd(stats + 2 * (int64_t)bb + 1, (double)t2);
          }
        }
        s1 = 0.f;
        s2 = 0.f;
      };
      for (int item = r_begin; item < r_end; ++item) {
        const int ls = item / H, oh = item - ls * H;
        const int b = step * spb + ls;
        if (b >= B) break;
        if (ls != cur_ls) {
          flush(cur_ls);
          cur_ls = ls;
          // issued now, consumed after the FMA block: the global-load latency hides behind the convolution
          add = bias_c + ((c_ok && cond) ? __ldg(cond + (int64_t)b * ldc + c) : 0.f);
        }
        const float* sx = sm + cur * slab + ls * HW * kDwCh;
        float acc[W];
#pragma unroll
        for (int i = 0; i < W; ++i) acc[i] = 0.f;
#pragma unroll
        for (int kh = 0; kh < 7; ++kh) {
          const int ih = oh + kh - 3;
          if (ih < 0 || ih >= H) continue;
          const float* row = sx + (ih * W) * kDwCh + cl;
#pragma unroll
          for (int iw = 0; iw < W; ++iw) {
            const float v = row[iw * kDwCh];
#pragma unroll
            for (int kw = 0; kw < 7; ++kw) {
              const int ow = iw - kw + 3;
              if (ow >= 0 && ow < W) acc[ow] = fmaf(v, wr[kh * 7 + kw], acc[ow]);
            }
          }
        }
        if (c_ok) {
          TOut* op = out + ((int64_t)b * HW + oh * W) * ldo + c;
#pragma unroll
          for (int i = 0; i < W; ++i) acc[i] += add;
          if (addend != nullptr) {
            const float* ap = addend + ((int64_t)b * HW + oh * W) * ldadd + c;
#pragma unroll
            for (int i = 0; i < W; ++i) acc[i] += ap[(int64_t)i * ldadd];
          }
#pragma unroll
          for (int i = 0; i < W; ++i) {
            op[(int64_t)i * ldo] = (TOut)acc[i];
            s1 += acc[i];
            s2 += acc[i] * acc[i];
          }
        }
      }
      flush(cur_ls);
    }
    __syncthreads();  // every warp is done with this buffer before the next prefetch overwrites it
    cur ^= 1;
  }
}

// ---- packed-fp32 variant (Blackwell FFMA2: fma.rn.f32x2 = two IEEE fp32 FMAs per instruction).
// The scalar kernel above is ISSUE-bound (ncu: 58 % of its instructions are FFMA, issue slots 71 % busy, FMA pipe 47 %).
// Here a lane owns a channel PAIR (x, weights and accumulators are 64-bit register pairs), so every multiply-add
// instruction does two, the x loads are 8-byte and the stores 4/8-byte.  A warp covers 16 channel pairs x 2 output rows:
// lanes 0-15 and 16-31 work on different rows (of two samples at the same row when the step holds several samples, so
// the two halves skip the same padding rows; rows oh, oh+1 of one sample at 16x16).  The 49 taps live in shared memory
// ([tap][32 channels]) and are re-read per kernel row (7 pair loads per 100 FFMA2).  Same accumulation order as the
// scalar kernel: results are bit-identical.
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

template <int W, typename TOut>
__global__ void __launch_bounds__(256, 3)
dwconv7_f2_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                  const float* __restrict__ bias, const float* __restrict__ cond, int64_t ldc,
                  TOut* __restrict__ out, int64_t ldo, double* __restrict__ stats, int B, int C, int H, int spb,
                  int flip, const float* __restrict__ addend, int64_t ldadd) {
  extern __shared__ __align__(16) float sm[];
  const int HW = H * W;
  const int slab = spb * HW * kDwCh;           // floats per buffer (spb samples per step)
  float* wsm = sm + 2 * slab;                  // [49][32] taps of this block's channels
  const int c0 = blockIdx.x * kDwCh;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int cp = lane & 15, half = lane >> 4;
  const int c = c0 + 2 * cp;
  const bool ok0 = c < C, ok1 = c + 1 < C;
  const int nsteps = (B + spb - 1) / spb;
  const int log_hw = 31 - __clz(HW);           // maps are powers of two

  auto prefetch = [&](int step, float* buf) {
    const int chunks = spb * HW * 8;
    for (int i = tid; i < chunks; i += 256) {
      const int q = i & 7, pix = i >> 3;
      const int b = step * spb + (pix >> log_hw);
      if (b < B && c0 + q * 4 < C)
        cp_async16(buf + pix * kDwCh + q * 4, x + ((int64_t)step * spb * HW + pix) * ldx + c0 + q * 4);
    }
    cp_async_commit();
  };

  for (int i = tid; i < 49 * kDwCh; i += 256) {
    const int tap = i >> 5, ch = i & 31;
    wsm[i] = (c0 + ch < C) ? __ldg(w + (int64_t)(c0 + ch) * 49 + (flip ? 48 - tap : tap)) : 0.f;
  }
  const float2 bias2 = make_float2((ok0 && bias) ? __ldg(bias + c) : 0.f, (ok1 && bias) ? __ldg(bias + c + 1) : 0.f);

  // work items of a step: (sample pair, row) when the step holds >= 2 samples, else row pairs of the one sample
  const bool by_sample = spb >= 2;
  const int items = by_sample ? ((spb + 1) >> 1) * H : (H + 1) >> 1;
  const int per_warp = (items + 7) >> 3;

  int step = blockIdx.y;
  if (step < nsteps) prefetch(step, sm);
  int cur = 0;
  for (; step < nsteps; step += gridDim.y) {
    const int nxt = step + gridDim.y;
    if (nxt < nsteps) {
      prefetch(nxt, sm + (cur ^ 1) * slab);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();   // (also orders the tap staging before the first use)
    {
      const int it_begin = warp * per_warp, it_end = min(items, it_begin + per_warp);
      float s1 = 0.f, s2 = 0.f;
      double d1 = 0.0, d2 = 0.0;
      int cur_key = -1, cur_ls = -1;
      // GroupNorm statistics: every ROW is summed the same way (fp32 over its channels and pixels, half-warp
      // reduction), the row totals are accumulated in fp64 and flushed once per sample -- so the numbers do not depend
      // on how the batch size groups rows into steps and warps, and a batch shard reproduces the unsharded result
      auto row_done = [&]() {     // executed by the whole warp (the item loop bounds are warp-uniform)
        float t1 = s1, t2 = s2;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          t1 += __shfl_xor_sync(0xffffffffu, t1, o);
          t2 += __shfl_xor_sync(0xffffffffu, t2, o);
        }
        d1 += (double)t1;
        d2 += (double)t2;
        s1 = 0.f;
        s2 = 0.f;
      };
      auto flush = [&]() {
        if (stats != nullptr && cur_ls >= 0 && cur_ls < spb && cp == 0) {
          const int bb = step * spb + cur_ls;
          if (bb < B) {
            atomicAdd(stats + 2 * (int64_t)bb, d1);
            atomicAdd(stats + 2 * (int64_t)bb + 1, d2);
          }
        }
        d1 = 0.0;
        d2 = 0.0;
      };
      for (int it = it_begin; it < it_end; ++it) {
        int ls, oh;
        if (by_sample) {
          const int sp2 = it / H;
          oh = it - sp2 * H;
          ls = 2 * sp2 + half;
        } else {
          ls = 0;
          oh = 2 * it + half;
        }
        const int key = by_sample ? it / H : 0;   // warp-uniform: the sample (pair) this item belongs to
        if (key != cur_key) {
          flush();
          cur_key = key;
          cur_ls = ls;
        }
        const int b = step * spb + ls;
        const bool row_ok = ls < spb && b < B && oh < H;
        float2 add2 = bias2;
        if (cond != nullptr && row_ok) {
          if (ok0) add2.x += __ldg(cond + (int64_t)b * ldc + c);
          if (ok1) add2.y += __ldg(cond + (int64_t)b * ldc + c + 1);
        }
        const float* sx = sm + cur * slab + ls * HW * kDwCh + 2 * cp;
        unsigned long long acc[W];
#pragma unroll
        for (int i = 0; i < W; ++i) acc[i] = 0ull;
#pragma unroll
        for (int kh = 0; kh < 7; ++kh) {
          const int ih = oh + kh - 3;
          if (ih < 0 || ih >= H || !row_ok) continue;
          unsigned long long wr[7];
#pragma unroll
          for (int kw = 0; kw < 7; ++kw)
            wr[kw] = *reinterpret_cast<const unsigned long long*>(wsm + (kh * 7 + kw) * kDwCh + 2 * cp);
          const float* row = sx + (ih * W) * kDwCh;
#pragma unroll
          for (int iw = 0; iw < W; ++iw) {
            const unsigned long long v = *reinterpret_cast<const unsigned long long*>(row + iw * kDwCh);
#pragma unroll
            for (int kw = 0; kw < 7; ++kw) {
              const int ow = iw - kw + 3;
              if (ow >= 0 && ow < W) acc[ow] = ffma2(v, wr[kw], acc[ow]);
            }
          }
        }
        if (row_ok && ok0) {
          const int64_t prow = (int64_t)b * HW + oh * W;
          TOut* op = out + prow * ldo + c;
          const float* ap = addend != nullptr ? addend + prow * ldadd + c : nullptr;
#pragma unroll
          for (int i = 0; i < W; ++i) {
            float2 r = unpack2(acc[i]);
            r.x += add2.x;
            r.y += add2.y;
            if (ap != nullptr) {
              if (ok1) {
                const float2 a2 = *reinterpret_cast<const float2*>(ap + (int64_t)i * ldadd);
                r.x += a2.x;
                r.y += a2.y;
              } else {
                r.x += ap[(int64_t)i * ldadd];
              }
            }
            if constexpr (sizeof(TOut) == 2) {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(r.x, r.y);
              if (ok1) *reinterpret_cast<__nv_bfloat162*>(op + (int64_t)i * ldo) = h2;
              else op[(int64_t)i * ldo] = __low2bfloat16(h2);
              r = __bfloat1622float2(h2);   // the statistics describe the tensor the next GEMM reads
            } else {
              if (ok1) *reinterpret_cast<float2*>(op + (int64_t)i * ldo) = r;
              else op[(int64_t)i * ldo] = r.x;
            }
            s1 += r.x;
            s2 = fmaf(r.x, r.x, s2);
            if (ok1) {
              s1 += r.y;
              s2 = fmaf(r.y, r.y, s2);
            }
          }
        }
        row_done();
      }
      flush();
    }
    __syncthreads();  // every warp is done with this buffer before the next prefetch overwrites it
    cur ^= 1;
  }
}

// ------------------------------------------------------------------------------ group statistics
// stats[b][g] += (sum, sumsq) over the pixels x channels of group g.  grid = (chunks, groups, B)
__global__ void __launch_bounds__(256)
group_stats_kernel(const void* __restrict__ x, int in_dtype, int64_t ldx, int HW, int C, int G,
                   double* __restrict__ stats) {
  __shared__ float red[32];
  const int g = blockIdx.y, b = blockIdx.z;
  const int cpg = C / G;
  const int64_t n = (int64_t)HW * cpg;
  float s1 = 0.f, s2 = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % cpg);
    const int64_t p = i / cpg;
    const int64_t off = ((int64_t)b * HW + p) * ldx + g * cpg + cc;
    const float v = in_dtype == SBM_F32 ? reinterpret_cast<const float*>(x)[off]
                                        : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[off]);
    s1 += v;
    s2 += v * v;
  }
  const float t1 = block_sum(s1, red);
  const float t2 = block_sum(s2, red);
  if (threadIdx.x == 0) {
    atomicAdd(stats + 2 * ((int64_t)b * G + g), (double)t1);
    atomicAdd(stats + 2 * ((int64_t)b * G + g) + 1, (double)t2);
  }
}

// Row-wise variant: a block walks whole pixel ROWS (all channels, 16-byte / 8-byte loads, coalesced) and keeps one
// partial sum per thread -- a thread's 4 channels always fall into one group (channels per group % 4 == 0).  The
// kernel above runs one block per (chunk, group, sample) over 16..128-byte pieces of every row with a division and a
// modulo per element: 107 us per launch = 0.74 TB/s on the GroupNorm32 layers of `UNetModel`, 38 % of that net's
// forward (launch list, round 2).  Partial sums meet in shared memory in a fixed order (deterministic per block).
template <typename TIn>
__global__ void __launch_bounds__(256)
group_stats_rows_kernel(const TIn* __restrict__ x, int64_t ldx, int HW, int C, int G, int pix_per_block,
                        double* __restrict__ stats) {
  __shared__ float red[2][8][64];            // [sum | sumsq][row of the block][group]
  const int tpr = C >> 2;                    // threads per pixel row
  const int rpb = 256 / tpr;                 // rows in flight per block
  const int q = threadIdx.x % tpr, r = threadIdx.x / tpr;
  const int b = blockIdx.y;
  const int cpg4 = (C / G) >> 2;             // threads per group within a row (1, 2, 4, 8 ...)
  float s1 = 0.f, s2 = 0.f;
  if (r < rpb) {
    const int p_end = min(HW, (int)(blockIdx.x + 1) * pix_per_block);
    for (int pix = blockIdx.x * pix_per_block + r; pix < p_end; pix += rpb) {
      const float4 v = load_qkv4(x + ((int64_t)b * HW + pix) * ldx + 4 * q);
      s1 += (v.x + v.y) + (v.z + v.w);
      s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
    }
  }
  // lanes of one group are consecutive (cpg4 <= 32 is a power of two: tpr is, G divides C)
  for (int o = 1; o < cpg4 && o < 32; o <<= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (r < rpb && (q % cpg4) == 0 && cpg4 <= 32) {
    red[0][r][q / cpg4] = s1;
    red[1][r][q / cpg4] = s2;
  }
  __syncthreads();
  if (threadIdx.x < 2 * G) {
    const int which = threadIdx.x / G, g = threadIdx.x % G;
    float t = 0.f;
    for (int rr = 0; rr < rpb; ++rr) t += red[which][rr][g];
    atomicAdd(stats + 2 * ((int64_t)b * G + g) + which, (double)t);
  }
}

// ------------------------------------------------------------------------------ GroupNorm apply
// y = act((x - mean)*rstd*gamma + beta) (+ residual).  grid = (chunks, B): a block streams a contiguous pixel range
// of ONE sample, so mean / rstd of its groups are derived once (fp64 -> fp32) into shared memory; 8 channels per
// thread (16-byte bf16 vectors / 2 x 16-byte fp32 vectors).
template <typename T>
__device__ __forceinline__ void load8(const T* p, float* v);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
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
__device__ __forceinline__ void store8(T* p, const float* v);
template <>
__device__ __forceinline__ void store8<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float* v) {
  uint32_t u[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    u[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(u[0], u[1], u[2], u[3]);
}

// A thread owns ONE channel octet (its gamma/beta/mean/rstd live in registers) and walks the pixels of the block's
// pixel range with 4 independent 16-byte loads in flight; consecutive threads = consecutive octets of one pixel.
// A block serves SB consecutive samples (SB > 1 when a sample is only a few hundred octets -- the 8x8 / 4x4 maps of the
// PolyMNIST net at 32k latents ran one block per sample for two pixels per thread: the fp64 mean / rstd prologue and
// the block turnover, not HBM, set the pace (1.4 TB/s, ncu); the statistics of all SB samples are reduced up front.
// MULTI = false is the one-sample-per-block kernel with the sample loop compiled away (the loop and the hoisted
// gamma / beta registers cost 30 registers and 15 % of the large-map launches when they were unconditional).
// MOD = false compiles the per-sample modulation / addend operands out (they cost 20 registers: 74 -> 94, i.e. the
// third resident block per SM and 15 % of the bandwidth of every plain GroupNorm-apply launch).
template <typename TIn, typename TOut, bool MULTI, bool MOD>
__global__ void __launch_bounds__(256)
groupnorm_apply_kernel(const TIn* __restrict__ x, int64_t ldx, const double* __restrict__ stats,
                       const float* __restrict__ gamma, const float* __restrict__ beta,
                       const float* __restrict__ residual, int64_t ldr, TOut* __restrict__ out, int64_t ldo,
                       float* __restrict__ out_f32, int64_t ldo_f32, int HW, int C, int G, float eps, int act,
                       int vec_ok, const float* __restrict__ mod_scale, const float* __restrict__ mod_shift,
                       int64_t ld_mod, const float* __restrict__ post_add, int64_t ld_post, int B, int SB) {
  __shared__ float s_mean[MULTI ? 256 : 64], s_rstd[MULTI ? 256 : 64];   // [SB][G], SB * G <= 256
  const int b_first = MULTI ? blockIdx.y * SB : blockIdx.y;
  const int nb = MULTI ? min(SB, B - b_first) : 1;
  const int cpg = C / G;
  if ((int)threadIdx.x < nb * G) {
    const double inv_n = 1.0 / ((double)HW * cpg);
    const int64_t slot = (int64_t)b_first * G + threadIdx.x;   // stats is [B][G][2]: the block's slots are contiguous
    const double s1 = stats[2 * slot], s2 = stats[2 * slot + 1];
    const double mean = s1 * inv_n;
    const double var = fmax(s2 * inv_n - mean * mean, 0.0);
    s_mean[threadIdx.x] = (float)mean;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int co = (C + 7) >> 3;  // channel octets per pixel
  const int tq = min(co, (int)blockDim.x);
  const int lanes = blockDim.x / tq;      // pixels processed side by side by one block
  const int pl = threadIdx.x / tq;
  if (pl >= lanes) return;
  const int pstride = lanes * gridDim.x;
  const int p0 = blockIdx.x * lanes + pl;
  for (int q = threadIdx.x - pl * tq; q < co; q += tq) {
    const int c = q * 8;
    const bool full = vec_ok && (c + 8 <= C);
    float gm[MULTI ? 8 : 1], bt[MULTI ? 8 : 1];   // gamma / beta of the octet, loaded once for the block's samples
    if constexpr (MULTI) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ce = min(c + e, C - 1);
        gm[e] = __ldg(gamma + ce);
        bt[e] = __ldg(beta + ce);
      }
    }
    for (int sb = 0; sb < nb; ++sb) {
      const int b = b_first + sb;
      float sc[8], sh[8], pa[MOD ? 8 : 1];  // y = act(x * sc + sh) + pa
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ce = min(c + e, C - 1);
        const int g = sb * G + ((G == 1) ? 0 : ce / cpg);
        float a = s_rstd[g] * (MULTI ? gm[MULTI ? e : 0] : __ldg(gamma + ce));
        float o = (MULTI ? bt[MULTI ? e : 0] : __ldg(beta + ce)) - s_mean[g] * a;
        if (MOD && mod_scale != nullptr) {  // per-sample modulation of the normalised value: n * (1 + scale) + shift
          const float m1 = 1.f + __ldg(mod_scale + (int64_t)b * ld_mod + ce);
          a *= m1;
          o = fmaf(o, m1, __ldg(mod_shift + (int64_t)b * ld_mod + ce));
        }
        sc[e] = a;
        sh[e] = o;
        if constexpr (MOD) pa[e] = (post_add != nullptr) ? __ldg(post_add + (int64_t)b * ld_post + ce) : 0.f;
      }
      auto one = [&](int pix_in_sample, const float* v_in) {
        const int64_t pix = (int64_t)b * HW + pix_in_sample;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float y = fmaf(v_in[e], sc[e], sh[e]);
          if (act == SBM_ACT_SILU) y = silu(y);
          else if (act == SBM_ACT_GELU) y = gelu_exact(y);
          v[e] = MOD ? y + pa[MOD ? e : 0] : y;
        }
        if (residual != nullptr) {
          const float* rp = residual + pix * ldr + c;
          if (full) {
            float r[8];
            load8<float>(rp, r);
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] += r[e];
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (c + e < C) v[e] += rp[e];
          }
        }
        if (out != nullptr) {
          TOut* op = out + pix * ldo + c;
          if (full) {
            store8<TOut>(op, v);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (c + e < C) op[e] = (TOut)v[e];
          }
        }
        if (out_f32 != nullptr) {
          float* op = out_f32 + pix * ldo_f32 + c;
          if (full) {
            store8<float>(op, v);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (c + e < C) op[e] = v[e];
          }
        }
      };
      int p = p0;
      if (full) {
        for (; p + 3 * pstride < HW; p += 4 * pstride) {
          float v[4][8];
#pragma unroll
          for (int u = 0; u < 4; ++u) load8<TIn>(x + ((int64_t)b * HW + p + u * pstride) * ldx + c, v[u]);
#pragma unroll
          for (int u = 0; u < 4; ++u) one(p + u * pstride, v[u]);
        }
      }
      for (; p < HW; p += pstride) {
        float v[8];
        const TIn* xp = x + ((int64_t)b * HW + p) * ldx + c;
        if (full) {
          load8<TIn>(xp, v);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = (c + e < C) ? (float)xp[e] : 0.f;
        }
        one(p, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------ time embedding
// mode 0: unet_model.py:40-47   [sin | cos], freq_k = exp(-k*log(1e4)/(half-1))
// mode 1: unet_openai.py:66-83  [cos | sin], freq_k = exp(-k*log(1e4)/half)
__global__ void time_embed_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out, float* out_f32,
                                  int B, int dim, int ld, int mode) {
  const int half = dim / 2;
  const int64_t total = (int64_t)B * ld;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % ld);
    const int b = (int)(idx / ld);
    float v = 0.f;
    if (j < 2 * half) {
      const int k = j < half ? j : j - half;
      const float denom = mode == 0 ? (float)(half - 1) : (float)half;
      // same fp32 evaluation order as the reference: exp(k * -(log(1e4)/denom)) then t * freq
      const float f = mode == 0 ? expf((float)k * -(9.210340371976184f / denom))
                                : expf(-9.210340371976184f * (float)k / denom);
      const float arg = t[b] * f;
      const bool first = j < half;
      v = (mode == 0) ? (first ? sinf(arg) : cosf(arg)) : (first ? cosf(arg) : sinf(arg));
    }
    out[idx] = __float2bfloat16_rn(v);
    if (out_f32 != nullptr) out_f32[idx] = v;
  }
}

// Small-image variant (the latent score nets: 3 x 16 x 16, 5 x 8 x 8): one block per sample stages the zero-padded
// image and a column -> window-offset table in shared memory, so a 16-byte store of 8 im2col columns costs 8 table and
// 8 data reads from shared memory instead of 8 x (2 divisions + 2 modulos + bounds tests + a global load): the generic
// kernel ran at 0.15-0.6 TB/s (ncu, round 1).
__global__ void __launch_bounds__(256)
stem_im2col_smem_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ a, int B, int C, int H, int W, int KH,
                        int KW, int ldk) {
  extern __shared__ float im_s[];                 // [C][H + KH - 1][W + KW - 1] zero-padded image, then the table
  const int HP = H + KH - 1, WP = W + KW - 1;
  const int K = C * KH * KW;
  const int groups = ldk >> 3;
  int* koff = reinterpret_cast<int*>(im_s + C * HP * WP);   // [groups * 8], -1 = padding column
  for (int k = threadIdx.x; k < groups * 8; k += blockDim.x) {
    int off = -1;
    if (k < K) {
      const int kw = k % KW, r = k / KW, kh = r % KH, c = r / KH;
      off = (c * HP + kh) * WP + kw;
    }
    koff[k] = off;
  }
  const int HWp = H * W;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < C * HP * WP; i += blockDim.x) {
      const int wp = i % WP, r = i / WP, hp = r % HP, c = r / HP;
      const int ih = hp - KH / 2, iw = wp - KW / 2;
      im_s[i] = (ih >= 0 && ih < H && iw >= 0 && iw < W) ? __ldg(x + ((int64_t)(b * C + c) * H + ih) * W + iw) : 0.f;
    }
    __syncthreads();
    __nv_bfloat16* ab = a + (int64_t)b * HWp * ldk;
    for (int idx = threadIdx.x; idx < HWp * groups; idx += blockDim.x) {
      const int gq = idx % groups, pin = idx / groups;
      const int h = pin / W, w = pin - h * W;
      const int base = h * WP + w;
      const int4 o0 = *reinterpret_cast<const int4*>(koff + gq * 8);
      const int4 o1 = *reinterpret_cast<const int4*>(koff + gq * 8 + 4);
      const int o[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
      uint32_t packed[4];
#pragma unroll
      for (int e2 = 0; e2 < 4; ++e2) {
        const float v0 = o[2 * e2] >= 0 ? im_s[o[2 * e2] + base] : 0.f;
        const float v1 = o[2 * e2 + 1] >= 0 ? im_s[o[2 * e2 + 1] + base] : 0.f;
        __nv_bfloat162 t = __floats2bfloat162_rn(v0, v1);
        packed[e2] = *reinterpret_cast<uint32_t*>(&t);
      }
      *reinterpret_cast<uint4*>(ab + (int64_t)pin * ldk + gq * 8) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    }
  }
}

// ------------------------------------------------------------------------------ linear attention core
// unet_model.py:162-177.  qkv: fp32 [B, n, ldq] with channels [q(h,d) | k(h,d) | v(h,d)], d = 32.
// One block per (head, sample).  out: bf16 [B, n, ldo], channel h*32+e.
constexpr int kHeadDim = 32;
__global__ void __launch_bounds__(256)
linear_attn_kernel(const float* __restrict__ qkv, int64_t ldq, __nv_bfloat16* __restrict__ out, int64_t ldo, int n,
                   int heads, float scale) {
  extern __shared__ float sm[];
  float* sq = sm;                          // [n][33]
  float* sk = sq + (size_t)n * 33;         // [n][33]
  float* sv = sk + (size_t)n * 33;         // [n][33]
  float* ctx = sv + (size_t)n * 33;        // [32][33]  ctx[d][e]
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int hid = heads * kHeadDim;
  {
    // stage q|k|v rows with 4 rows (12 independent 128-byte requests per warp) in flight
    const float* base = qkv + (int64_t)b * n * ldq + h * kHeadDim + lane;
    int p = warp;
    for (; p + 3 * nwarp < n; p += 4 * nwarp) {
      float r[12];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* row = base + (int64_t)(p + u * nwarp) * ldq;
        r[3 * u] = __ldg(row);
        r[3 * u + 1] = __ldg(row + hid);
        r[3 * u + 2] = __ldg(row + 2 * hid);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int pp = p + u * nwarp;
        sq[pp * 33 + lane] = r[3 * u];
        sk[pp * 33 + lane] = r[3 * u + 1];
        sv[pp * 33 + lane] = r[3 * u + 2];
      }
    }
    for (; p < n; p += nwarp) {
      const float* row = base + (int64_t)p * ldq;
      sq[p * 33 + lane] = __ldg(row);
      sk[p * 33 + lane] = __ldg(row + hid);
      sv[p * 33 + lane] = __ldg(row + 2 * hid);
    }
  }
  __syncwarp();
  for (int p = warp; p < n; p += nwarp) {
    // q: softmax over d (the 32 lanes), then * scale
    const float qv = sq[p * 33 + lane];
    float m = qv;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float e = expf(qv - m);
    const float s = warp_sum(e);
    sq[p * 33 + lane] = e / s * scale;
  }
  __syncthreads();
  // k: softmax over the n positions, per channel d
  for (int d = warp; d < kHeadDim; d += nwarp) {
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
  // context[d][e] = sum_n k[d][n] v[e][n]
  for (int i = tid; i < kHeadDim * kHeadDim; i += blockDim.x) {
    const int d = i >> 5, e = i & 31;
    float acc = 0.f;
    for (int p = 0; p < n; ++p) acc = fmaf(sk[p * 33 + d], sv[p * 33 + e], acc);
    ctx[d * 33 + e] = acc;
  }
  __syncthreads();
  // out[e][n] = sum_d context[d][e] q[d][n]
  for (int p = warp; p < n; p += nwarp) {
    float acc = 0.f;
#pragma unroll 8
    for (int d = 0; d < kHeadDim; ++d) acc = fmaf(ctx[d * 33 + lane], sq[p * 33 + d], acc);
    out[((int64_t)b * n + p) * ldo + h * kHeadDim + lane] = __float2bfloat16_rn(acc);
  }
}

// Register-tiled variant (the one the forward pass uses).  Same math, organised so that global memory is read with
// ONE wave of cp.async requests per block and every shared-memory access of the two small matmuls is a conflict-free
// 16-byte load feeding 16 FMAs:
//   load : raw q | k | v rows -> sq, sk, sv [p][32] with cp.async (all requests in flight at once)
//   q    : soft-max over d in place (lane = channel, warp shuffles); k: column max with lane = channel
//   k    : exp(k - max) in place + column sums; the 1/sum factor is applied to the 32x32 context, not to n x 32
//   ctx  : thread = 4(d) x 4(e) tile, the n positions split over 4 thread groups, partial sums reduced in smem
//   out  : thread = 4(p) x 4(e) tile, out[p][e] = sum_d ctx[d][e] q[p][d]
__global__ void __launch_bounds__(256, 2)
linear_attn_tiled_kernel(const float* __restrict__ qkv, int64_t ldq, __nv_bfloat16* __restrict__ out, int64_t ldo, int n,
                         int heads, float scale) {
  extern __shared__ __align__(16) float sm[];
  const int n4 = (n + 3) & ~3;
  float* sq = sm;                           // [n4][32]
  float* sk = sq + n4 * 32;                 // [n4][32]   (re-used for the 4 partial contexts [4][32][32])
  float* sv = sk + max(n4 * 32, 4 * 1024);  // [n4][32]
  float* ctx = sv + n4 * 32;                // [32][32]
  float* red = ctx + 1024;                  // [2][8][32]
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hid = heads * kHeadDim;
  const float* base = qkv + (int64_t)b * n * ldq + h * kHeadDim;

  // ---- load: 3 matrices x n rows x 8 sixteen-byte chunks.  q rows are stored with their 16-byte chunks XOR-swizzled
  // by (row & 7) so that a thread can later read "its" whole row with conflict-free 16-byte loads.
  {  // thread = (16-byte chunk q8, row p0 + 32 j): pointer increments only (the flat-index version spent 22 % of the
     // kernel's instructions on the div / mod of this loop)
    const int q8 = tid & 7, p0 = tid >> 3;
    const float* src = base + (int64_t)p0 * ldq + q8 * 4;
    const int64_t rstep = 32 * ldq;
    for (int p = p0; p < n; p += 32, src += rstep) {
      cp_async16(sq + p * 32 + ((q8 ^ (p & 7)) << 2), src);
      cp_async16(sk + p * 32 + q8 * 4, src + hid);
      cp_async16(sv + p * 32 + q8 * 4, src + 2 * hid);
    }
  }
  cp_async_commit();
  for (int i = n * 32 + tid; i < n4 * 32; i += 256) {  // padded rows: q = v = 0, k = -inf
    sq[i] = 0.f;
    sk[i] = -INFINITY;
    sv[i] = 0.f;
  }
  cp_async_wait<0>();
  __syncthreads();
  // ---- q soft-max over d: one thread = one row, all 32 channels in registers (no shuffle chains)
  for (int p = tid; p < n; p += 256) {
    float4* rowp = reinterpret_cast<float4*>(sq + p * 32);
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = rowp[k ^ (p & 7)];
    float m = v[0].x;
#pragma unroll
    for (int k = 0; k < 8; ++k) m = fmaxf(fmaxf(fmaxf(m, v[k].x), fmaxf(v[k].y, v[k].z)), v[k].w);
    float ssum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k].x = __expf(v[k].x - m); v[k].y = __expf(v[k].y - m); v[k].z = __expf(v[k].z - m); v[k].w = __expf(v[k].w - m);
      ssum += (v[k].x + v[k].y) + (v[k].z + v[k].w);
    }
    const float inv = scale / ssum;
#pragma unroll
    for (int k = 0; k < 8; ++k) rowp[k ^ (p & 7)] = make_float4(v[k].x * inv, v[k].y * inv, v[k].z * inv, v[k].w * inv);
  }
  // ---- k column max with lane = channel
  float kmax = -INFINITY;
  for (int p = warp; p < n; p += 8) kmax = fmaxf(kmax, sk[p * 32 + lane]);
  red[warp * 32 + lane] = kmax;
  __syncthreads();
  float cmax = red[lane];
#pragma unroll
  for (int w2 = 1; w2 < 8; ++w2) cmax = fmaxf(cmax, red[w2 * 32 + lane]);
  // ---- k: exp(k - max) in place, column sums
  float ksum = 0.f;
  for (int p = warp; p < n4; p += 8) {
    const float e = __expf(sk[p * 32 + lane] - cmax);  // padded rows hold -inf -> 0
    sk[p * 32 + lane] = e;
    ksum += e;
  }
  red[256 + warp * 32 + lane] = ksum;
  __syncthreads();
  // ---- context partials: thread = (group g, 4x4 tile (d0, e0))
  {
    const int g = tid >> 6, t = tid & 63;
    const int d0 = (t >> 3) * 4, e0 = (t & 7) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j2 = 0; j2 < 4; ++j2) acc[i][j2] = 0.f;
    const int per = (n4 + 3) >> 2;
    const int pe = min(n4, (g + 1) * per);
    for (int p = g * per; p < pe; ++p) {
      const float4 kd = *reinterpret_cast<const float4*>(sk + p * 32 + d0);
      const float4 ve = *reinterpret_cast<const float4*>(sv + p * 32 + e0);
      const float kk[4] = {kd.x, kd.y, kd.z, kd.w}, vv[4] = {ve.x, ve.y, ve.z, ve.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) acc[i][j2] = fmaf(kk[i], vv[j2], acc[i][j2]);
    }
    __syncthreads();  // everyone is done reading sk before it is overwritten with the partial contexts
    float* part = sk + g * 1024;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      *reinterpret_cast<float4*>(part + (d0 + i) * 32 + e0) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  __syncthreads();
  for (int i = tid; i < 1024; i += 256) {
    const int d = i >> 5;
    float ssum = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < 8; ++w2) ssum += red[256 + w2 * 32 + d];
    ctx[i] = (sk[i] + sk[1024 + i] + sk[2048 + i] + sk[3072 + i]) / ssum;
  }
  __syncthreads();
  // ---- out: thread = 4(p) x 4(e) tile; the 8 lanes of a quarter warp share p0, so the q loads broadcast
  const int ptiles = n4 >> 2;
  for (int t = tid; t < ptiles * 8; t += 256) {
    const int p0 = (t >> 3) * 4, e0 = (t & 7) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j2 = 0; j2 < 4; ++j2) acc[i][j2] = 0.f;
#pragma unroll 2
    for (int d4 = 0; d4 < 8; ++d4) {
      float qq[4][4], cc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t4 = *reinterpret_cast<const float4*>(sq + (p0 + i) * 32 + ((d4 ^ ((p0 + i) & 7)) << 2));
        qq[i][0] = t4.x; qq[i][1] = t4.y; qq[i][2] = t4.z; qq[i][3] = t4.w;
        const float4 c4 = *reinterpret_cast<const float4*>(ctx + (d4 * 4 + i) * 32 + e0);
        cc[i][0] = c4.x; cc[i][1] = c4.y; cc[i][2] = c4.z; cc[i][3] = c4.w;
      }
#pragma unroll
      for (int dd = 0; dd < 4; ++dd)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j2 = 0; j2 < 4; ++j2) acc[i][j2] = fmaf(qq[i][dd], cc[dd][j2], acc[i][j2]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (p0 + i >= n) break;
      __nv_bfloat162 lo = __floats2bfloat162_rn(acc[i][0], acc[i][1]), hi = __floats2bfloat162_rn(acc[i][2], acc[i][3]);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(out + ((int64_t)b * n + p0 + i) * ldo + h * kHeadDim + e0) = pk;
    }
  }
}

// Tensor-core variant (n = 64 or 256 positions, the 8x8 and 16x16 levels that carry ~90 % of the attention time).
// The register-tiled kernel above is FMA-issue bound (~5 k instructions per thread, 1.2 TB/s at 16x16: ncu, round 1);
// here the two 32-wide products run as bf16 mma.sync.m16n8k16 with fp32 accumulation:
//   load : every thread reads 16-byte chunks of R = n/32 rows of q, k and v straight into registers (no staging)
//   q    : soft-max over d across the 8 lanes that share a row (3 xor-shuffles), * scale, -> bf16 qb[p][d]
//   k    : column max / exp / column sums from the registers (lanes with equal tid & 7 share channels: 2 shuffles +
//          one shared-memory exchange between the warps), exp(k - max) -> bf16 kb[p][d]; 1/sum goes into the context
//   v    : -> bf16 vb[p][e]
//   ctx  : warp = one 16x8 tile of ctx[d][e] = sum_p kb[p][d] vb[p][e]   (A, B via ldmatrix.trans), * 1/ksum -> cb
//   out  : warp = two 16-row tiles of out[p][e] = sum_d qb[p][d] cb[d][e]; the tile is written back over the warp's own
//          qb rows and leaves with 16-byte coalesced stores
// Rows are 64 bytes; 16-byte chunk c of row r lives at chunk c ^ ((r >> 1) & 3), which makes every ldmatrix phase
// (8 rows x 16 bytes) conflict-free.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t swz_off(int row, int col) {  // byte offset of element (row, col) of a [*][32] bf16 tile
  return (uint32_t)(row * 64 + ((((col >> 3) ^ (row >> 1)) & 3) << 4) + ((col & 7) << 1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int R, typename TIn>  // rows per thread: n = 32 * R
__global__ void __launch_bounds__(256, 2)
linear_attn_mma_kernel(const TIn* __restrict__ qkv, int64_t ldq, __nv_bfloat16* __restrict__ out, int64_t ldo,
                       int heads, float scale) {
  constexpr int n = 32 * R;
  extern __shared__ __align__(128) uint8_t attn_smem[];
  uint8_t* qb = attn_smem;                   // [n][32] bf16, swizzled
  uint8_t* kb = qb + n * 64;
  uint8_t* vb = kb + n * 64;
  uint8_t* cb = vb + n * 64;                 // [32][32] bf16 context
  float (*red)[8][32] = reinterpret_cast<float (*)[8][32]>(cb + 32 * 64);   // [2][8][32] per-warp column maxima / sums
  float* kinv = reinterpret_cast<float*>(cb + 32 * 64 + 2 * 8 * 32 * 4);
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hid = heads * kHeadDim;
  const int q8 = tid & 7, p0 = tid >> 3;        // 16-byte chunk (channels 4 q8 .. 4 q8 + 3) of rows p0 + 32 j
  const TIn* src = qkv + ((int64_t)b * n + p0) * ldq + h * kHeadDim + q8 * 4;

  // ---- load: 3 R independent 16-byte (fp32) / 8-byte (bf16) loads per thread
  float4 qv[R], kv[R], vv[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const TIn* s = src + (int64_t)(32 * j) * ldq;
    qv[j] = load_qkv4(s);
    kv[j] = load_qkv4(s + hid);
    vv[j] = load_qkv4(s + 2 * hid);
  }
  // ---- v -> bf16; q soft-max over d (8 lanes share a row); k column max over this thread's rows
  float4 kmax4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int p = p0 + 32 * j;
    *reinterpret_cast<uint2*>(vb + swz_off(p, q8 * 4)) =
        make_uint2(pack_bf16x2(vv[j].x, vv[j].y), pack_bf16x2(vv[j].z, vv[j].w));
    float m = fmaxf(fmaxf(qv[j].x, qv[j].y), fmaxf(qv[j].z, qv[j].w));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
    const float e0 = __expf(qv[j].x - m), e1 = __expf(qv[j].y - m), e2 = __expf(qv[j].z - m), e3 = __expf(qv[j].w - m);
    float ssum = (e0 + e1) + (e2 + e3);
    ssum += __shfl_xor_sync(0xffffffffu, ssum, 1);
    ssum += __shfl_xor_sync(0xffffffffu, ssum, 2);
    ssum += __shfl_xor_sync(0xffffffffu, ssum, 4);
    const float inv = scale / ssum;
    *reinterpret_cast<uint2*>(qb + swz_off(p, q8 * 4)) =
        make_uint2(pack_bf16x2(e0 * inv, e1 * inv), pack_bf16x2(e2 * inv, e3 * inv));
    kmax4.x = fmaxf(kmax4.x, kv[j].x); kmax4.y = fmaxf(kmax4.y, kv[j].y);
    kmax4.z = fmaxf(kmax4.z, kv[j].z); kmax4.w = fmaxf(kmax4.w, kv[j].w);
  }
  // lanes q8, q8 + 8, q8 + 16, q8 + 24 hold the same channels
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    kmax4.x = fmaxf(kmax4.x, __shfl_xor_sync(0xffffffffu, kmax4.x, o));
    kmax4.y = fmaxf(kmax4.y, __shfl_xor_sync(0xffffffffu, kmax4.y, o));
    kmax4.z = fmaxf(kmax4.z, __shfl_xor_sync(0xffffffffu, kmax4.z, o));
    kmax4.w = fmaxf(kmax4.w, __shfl_xor_sync(0xffffffffu, kmax4.w, o));
  }
  if (lane < 8) *reinterpret_cast<float4*>(&red[0][warp][q8 * 4]) = kmax4;
  __syncthreads();
  float4 cmax = *reinterpret_cast<const float4*>(&red[0][0][q8 * 4]);
#pragma unroll
  for (int w2 = 1; w2 < 8; ++w2) {
    const float4 t = *reinterpret_cast<const float4*>(&red[0][w2][q8 * 4]);
    cmax.x = fmaxf(cmax.x, t.x); cmax.y = fmaxf(cmax.y, t.y); cmax.z = fmaxf(cmax.z, t.z); cmax.w = fmaxf(cmax.w, t.w);
  }
  // ---- k: exp(k - max) -> bf16, column sums
  float4 ks = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int p = p0 + 32 * j;
    const float e0 = __expf(kv[j].x - cmax.x), e1 = __expf(kv[j].y - cmax.y), e2 = __expf(kv[j].z - cmax.z),
                e3 = __expf(kv[j].w - cmax.w);
    ks.x += e0; ks.y += e1; ks.z += e2; ks.w += e3;
    *reinterpret_cast<uint2*>(kb + swz_off(p, q8 * 4)) = make_uint2(pack_bf16x2(e0, e1), pack_bf16x2(e2, e3));
  }
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    ks.x += __shfl_xor_sync(0xffffffffu, ks.x, o); ks.y += __shfl_xor_sync(0xffffffffu, ks.y, o);
    ks.z += __shfl_xor_sync(0xffffffffu, ks.z, o); ks.w += __shfl_xor_sync(0xffffffffu, ks.w, o);
  }
  if (lane < 8) *reinterpret_cast<float4*>(&red[1][warp][q8 * 4]) = ks;
  __syncthreads();   // qb, kb, vb and the per-warp sums are complete
  if (tid < 32) {
    float t = 0.f;
#pragma unroll
    for (int w2 = 0; w2 < 8; ++w2) t += red[1][w2][tid];
    kinv[tid] = 1.f / t;
  }
  // ---- context: warp = tile (mi, ni) of ctx[d][e], d in [16 mi, +16), e in [8 ni, +8); K = all n positions
  {
    const int mi = warp >> 2, ni = warp & 3;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    const int jm = lane >> 3, r = lane & 7;
    const uint32_t kb_s = smem_addr(kb), vb_s = smem_addr(vb);
#pragma unroll 4
    for (int ks0 = 0; ks0 < n; ks0 += 16) {
      uint32_t a[4], bfr[2];
      // A[m = d][k = p] = kb[p][d] (stored [k][m]): matrices (k 0-7, m 0-7), (k 0-7, m 8-15), (k 8-15, m 0-7), (k 8-15, m 8-15)
      ldsm_x4_t(kb_s + swz_off(ks0 + r + ((jm & 2) ? 8 : 0), 16 * mi + ((jm & 1) ? 8 : 0)), a);
      // B[k = p][n = e] = vb[p][e] (stored [k][n]): matrices (k 0-7), (k 8-15) at columns 8 ni
      ldsm_x2_t(vb_s + swz_off(ks0 + r + ((jm & 1) ? 8 : 0), 8 * ni), bfr);
      mma_bf16_16816(c, a, bfr);
    }
    __syncthreads();   // kinv is visible (and every warp is past its reads of red[1])
    const int g = lane >> 2, t2 = (lane & 3) * 2;
    const int d_lo = 16 * mi + g, d_hi = d_lo + 8, e = 8 * ni + t2;
    *reinterpret_cast<uint32_t*>(cb + swz_off(d_lo, e)) = pack_bf16x2(c[0] * kinv[d_lo], c[1] * kinv[d_lo]);
    *reinterpret_cast<uint32_t*>(cb + swz_off(d_hi, e)) = pack_bf16x2(c[2] * kinv[d_hi], c[3] * kinv[d_hi]);
  }
  __syncthreads();
  // ---- out: warp = 16-row tiles mt = warp, warp + 8, ...; B fragments of the whole 32x32 context loaded once
  {
    const int jm = lane >> 3, r = lane & 7;
    const uint32_t qb_s = smem_addr(qb), cb_s = smem_addr(cb);
    uint32_t bf[2][4][2];
#pragma unroll
    for (int kk = 0; kk < 2; ++kk)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) ldsm_x2_t(cb_s + swz_off(16 * kk + r + ((jm & 1) ? 8 : 0), 8 * ni), bf[kk][ni]);
    const int g = lane >> 2, t2 = (lane & 3) * 2;
    for (int mt = warp; mt < n / 16; mt += 8) {
      float c[4][4];
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) { c[ni][0] = 0.f; c[ni][1] = 0.f; c[ni][2] = 0.f; c[ni][3] = 0.f; }
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t a[4];
        // A[m = p][k = d] = qb[p][d] (row-major): matrices (m 0-7, k 0-7), (m 8-15, k 0-7), (m 0-7, k 8-15), (m 8-15, k 8-15)
        ldsm_x4(qb_s + swz_off(16 * mt + r + ((jm & 1) ? 8 : 0), 16 * kk + ((jm & 2) ? 8 : 0)), a);
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) mma_bf16_16816(c[ni], a, bf[kk][ni]);
      }
      __syncwarp();   // all lanes have read this tile's qb rows: overwrite them with the output tile
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        *reinterpret_cast<uint32_t*>(qb + swz_off(16 * mt + g, 8 * ni + t2)) = pack_bf16x2(c[ni][0], c[ni][1]);
        *reinterpret_cast<uint32_t*>(qb + swz_off(16 * mt + g + 8, 8 * ni + t2)) = pack_bf16x2(c[ni][2], c[ni][3]);
      }
      __syncwarp();
      // 16 rows x 4 chunks of 16 bytes: two per lane, 4 consecutive lanes cover one 64-byte output row
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = lane + 32 * i, row = 16 * mt + (idx >> 2), ch = idx & 3;
        const uint4 v = *reinterpret_cast<const uint4*>(qb + swz_off(row, ch * 8));
        *reinterpret_cast<uint4*>(out + ((int64_t)b * n + row) * ldo + h * kHeadDim + ch * 8) = v;
      }
    }
  }
}

// Small maps (n <= 16 positions: the 4x4, 2x2 and 1x1 levels).  One WARP per (head, sample), lane = channel; the
// products broadcast their scalar operand with warp shuffles, nothing goes through shared memory.  The block-per-
// (head, sample) kernels above spend 35-41 us per launch here on 4096 nearly empty blocks (ncu, round 2).
template <int NMAX>
__global__ void __launch_bounds__(256)
linear_attn_small_kernel(const float* __restrict__ qkv, int64_t ldq, __nv_bfloat16* __restrict__ out, int64_t ldo,
                         int n, int heads, int pairs, float scale) {
  const int lane = threadIdx.x & 31;
  const int hb = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (hb >= pairs) return;
  const int b = hb / heads, h = hb - b * heads;
  const int hid = heads * kHeadDim;
  const float* base = qkv + (int64_t)b * n * ldq + h * kHeadDim + lane;
  float q[NMAX], k[NMAX], v[NMAX];
#pragma unroll
  for (int p = 0; p < NMAX; ++p) {
    const bool on = p < n;
    q[p] = on ? __ldg(base + (int64_t)p * ldq) : 0.f;
    k[p] = on ? __ldg(base + (int64_t)p * ldq + hid) : -INFINITY;
    v[p] = on ? __ldg(base + (int64_t)p * ldq + 2 * hid) : 0.f;
  }
  // q: soft-max over the 32 channels of a row (across lanes), * scale
#pragma unroll
  for (int p = 0; p < NMAX; ++p) {
    if (p >= n) break;
    float m = q[p];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float e = __expf(q[p] - m);
    q[p] = e * (scale / warp_sum(e));
  }
  // k: soft-max over the positions of a channel (inside the lane)
  float km = k[0];
#pragma unroll
  for (int p = 1; p < NMAX; ++p) km = fmaxf(km, k[p]);
  float ksum = 0.f;
#pragma unroll
  for (int p = 0; p < NMAX; ++p) {
    k[p] = __expf(k[p] - km);   // padded rows hold -inf -> 0
    ksum += k[p];
  }
  const float kinv = 1.f / ksum;
  // the two products need a SCALAR of another lane per multiply-add (k~[p][d], q~[p][d] live in lane d): stage both in
  // the warp's shared-memory rows and read them back as 16-byte broadcasts (a shuffle per multiply-add made the kernel
  // SHFL-bound: 1 warp shuffle per clock and SM, 34 us at n = 16)
  __shared__ __align__(16) float stage[8][2][NMAX][kHeadDim];
  float (*sq)[kHeadDim] = stage[threadIdx.x >> 5][0];
  float (*sk)[kHeadDim] = stage[threadIdx.x >> 5][1];
#pragma unroll
  for (int p = 0; p < NMAX; ++p) {
    sq[p][lane] = q[p];
    sk[p][lane] = k[p] * kinv;
  }
  __syncwarp();
  // context[d][e] (lane = e)
  float ctx[kHeadDim];
#pragma unroll
  for (int d = 0; d < kHeadDim; ++d) ctx[d] = 0.f;
#pragma unroll
  for (int p = 0; p < NMAX; ++p) {
    if (p >= n) break;
#pragma unroll
    for (int d4 = 0; d4 < kHeadDim; d4 += 4) {
      const float4 kk = *reinterpret_cast<const float4*>(&sk[p][d4]);
      ctx[d4] = fmaf(kk.x, v[p], ctx[d4]);
      ctx[d4 + 1] = fmaf(kk.y, v[p], ctx[d4 + 1]);
      ctx[d4 + 2] = fmaf(kk.z, v[p], ctx[d4 + 2]);
      ctx[d4 + 3] = fmaf(kk.w, v[p], ctx[d4 + 3]);
    }
  }
  // out[p][e] = sum_d ctx[d][e] q~[p][d]
  for (int p = 0; p < n; ++p) {
    float acc = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < kHeadDim; d4 += 4) {
      const float4 qq = *reinterpret_cast<const float4*>(&sq[p][d4]);
      acc = fmaf(ctx[d4], qq.x, acc);
      acc = fmaf(ctx[d4 + 1], qq.y, acc);
      acc = fmaf(ctx[d4 + 2], qq.z, acc);
      acc = fmaf(ctx[d4 + 3], qq.w, acc);
    }
    out[((int64_t)b * n + p) * ldo + h * kHeadDim + lane] = __float2bfloat16_rn(acc);
  }
}

// ------------------------------------------------------------------------------ softmax attention core
// qkv: fp32 [B, n, ldq]; layout selected by (q_off, k_off, v_off, head_stride): channel of (head, d) for q is
// q_off + head*head_stride + d.   out[b, i, o_off + head*dh + d] = sum_j softmax_j(scale * q_i . k_j) v_j[d]
//   unet_model.py:135-149  : q_off=0, k_off=hid, v_off=2*hid, head_stride=dh, scale=dh^-0.5
//   unet_openai.py:333-358 : per head [q|k|v] blocks: q_off=0,k_off=ch,v_off=2ch, head_stride=3ch, scale=ch^-0.5
// One block per (head, sample); K/V rows stream through shared memory in chunks of 32 channels.
__global__ void __launch_bounds__(256)
softmax_attn_kernel(const float* __restrict__ qkv, int64_t ldq, __nv_bfloat16* __restrict__ out, int64_t ldo, int n,
                    int dh, int q_off, int k_off, int v_off, int head_stride, float scale) {
  extern __shared__ float sm[];
  float* sS = sm;                      // [n][n+1] scores / probabilities
  float* sA = sS + (size_t)n * (n + 1);  // [n][33] chunk of q or v
  float* sB = sA + (size_t)n * 33;       // [n][33] chunk of k
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const float* base = qkv + (int64_t)b * n * ldq + h * head_stride;
  for (int i = tid; i < n * (n + 1); i += blockDim.x) sS[i] = 0.f;
  for (int d0 = 0; d0 < dh; d0 += 32) {
    __syncthreads();
    for (int p = warp; p < n; p += nwarp) {
      const bool ok = d0 + lane < dh;
      sA[p * 33 + lane] = ok ? base[(int64_t)p * ldq + q_off + d0 + lane] : 0.f;
      sB[p * 33 + lane] = ok ? base[(int64_t)p * ldq + k_off + d0 + lane] : 0.f;
    }
    __syncthreads();
    for (int idx = tid; idx < n * n; idx += blockDim.x) {
      const int i = idx / n, j = idx - i * n;
      float acc = 0.f;
#pragma unroll 8
      for (int d = 0; d < 32; ++d) acc = fmaf(sA[i * 33 + d], sB[j * 33 + d], acc);
      sS[i * (n + 1) + j] += acc;
    }
  }
  __syncthreads();
  for (int i = warp; i < n; i += nwarp) {
    float m = -INFINITY;
    for (int j = lane; j < n; j += 32) m = fmaxf(m, sS[i * (n + 1) + j] * scale);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int j = lane; j < n; j += 32) {
      const float e = expf(sS[i * (n + 1) + j] * scale - m);
      sS[i * (n + 1) + j] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.f / s;
    for (int j = lane; j < n; j += 32) sS[i * (n + 1) + j] *= inv;
  }
  for (int d0 = 0; d0 < dh; d0 += 32) {
    __syncthreads();
    for (int p = warp; p < n; p += nwarp)
      sA[p * 33 + lane] = (d0 + lane < dh) ? base[(int64_t)p * ldq + v_off + d0 + lane] : 0.f;
    __syncthreads();
    for (int i = warp; i < n; i += nwarp) {
      float acc = 0.f;
      for (int j = 0; j < n; ++j) acc = fmaf(sS[i * (n + 1) + j], sA[j * 33 + lane], acc);
      if (d0 + lane < dh) out[((int64_t)b * n + i) * ldo + h * dh + d0 + lane] = __float2bfloat16_rn(acc);
    }
  }
}

// nearest 2x upsampling, bf16 channels-last, 8 channels (16 bytes) per thread
__global__ void upsample2x_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out,
                                  int64_t ldo, int B, int H, int W, int C8) {
  const int64_t total = (int64_t)B * (2 * H) * (2 * W) * C8;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(idx % C8);
    const int64_t pix = idx / C8;
    const int ow = (int)(pix % (2 * W));
    const int oh = (int)((pix / (2 * W)) % (2 * H));
    const int b = (int)(pix / ((int64_t)4 * W * H));
    const uint4 v = *reinterpret_cast<const uint4*>(x + (((int64_t)b * H + (oh >> 1)) * W + (ow >> 1)) * ldx + q * 8);
    *reinterpret_cast<uint4*>(out + pix * ldo + q * 8) = v;
  }
}

static int grid_for(int64_t total, int threads) {
  const int64_t want = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)std::max<int64_t>(1, std::min(want, cap));
}

template <int W, typename TOut>
static int dwconv7_pipe_launch(const float* x, int64_t ldx, const float* w, const float* bias, const float* cond,
                               int64_t ldc, void* out, int64_t ldo, double* stats, int B, int H, int C, int flip,
                               const float* addend, int64_t ldadd, cudaStream_t st) {
  // 256 pixels (32 KB) per step: 1 sample of 16x16, 4 of 8x8, 16 of 4x4, ...
  const int spb = std::max(1, std::min(B, 256 / (H * W)));
  // packed-fp32 kernel: needs even strides and 8-byte aligned rows for its pair loads / stores
  static const bool f2_env = [] { const char* e = getenv("SBM_DWCONV_F32X2"); return e ? atoi(e) != 0 : true; }();
  const auto al = [](const void* ptr, int64_t ld, int bytes) {
    return ptr == nullptr || ((ld % 2 == 0) && (reinterpret_cast<uintptr_t>(ptr) % bytes == 0));
  };
  const bool f2 = f2_env && al(out, ldo, 2 * (int)sizeof(TOut)) && al(addend, ldadd, 8);
  const size_t smem = (size_t)2 * spb * H * W * kDwCh * sizeof(float) + (f2 ? 49 * kDwCh * sizeof(float) : 0);
  static size_t configured = 0, configured_occ = 0;
  if (smem > 48 * 1024 && smem > configured) {
    SBM_CUDA_OK(cudaFuncSetAttribute(dwconv7_pipe_kernel<W, TOut>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    SBM_CUDA_OK(cudaFuncSetAttribute(dwconv7_f2_kernel<W, TOut>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    configured = smem;
  }
  const int chunks = (C + kDwCh - 1) / kDwCh;
  const int nsteps = (B + spb - 1) / spb;
  // exactly ONE wave of resident blocks (2-3 per SM), each walking several steps so the prefetch overlaps: the grid
  // is rounded DOWN to the resident slots -- a handful of blocks spilling into a second wave would run their whole
  // share of the steps after everyone else has finished (measured: 300 blocks on 296 slots cost 1.5x)
  static int per_sm = 0, per_sm_f2 = 0;
  if (per_sm == 0 || smem > configured_occ) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dwconv7_pipe_kernel<W, TOut>, 256, smem) != cudaSuccess ||
        per_sm <= 0)
      per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_f2, dwconv7_f2_kernel<W, TOut>, 256, smem) != cudaSuccess ||
        per_sm_f2 <= 0)
      per_sm_f2 = 1;
    configured_occ = smem;
  }
  int gy = std::max(1, std::min(nsteps, (f2 ? per_sm_f2 : per_sm) * sm_count() / chunks));
  dim3 grid(chunks, gy);
  if (f2)
    dwconv7_f2_kernel<W, TOut><<<grid, 256, smem, st>>>(x, ldx, w, bias, cond, ldc, (TOut*)out, ldo, stats, B, C, H, spb,
                                                        flip, addend, ldadd);
  else
    dwconv7_pipe_kernel<W, TOut><<<grid, 256, smem, st>>>(x, ldx, w, bias, cond, ldc, (TOut*)out, ldo, stats, B, C, H,
                                                          spb, flip, addend, ldadd);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int dwconv7_mma_launch(const float* x, int64_t ldx, const float* w, const float* bias, const float* cond, int64_t ldc,
                       void* out_bf16, int64_t ldo, double* stats, int B, int H, int W, int C, cudaStream_t st);

static int dwconv7_launch(const float* x, int64_t ldx, const float* w, const float* bias, const float* cond,
                          int64_t ldc, void* out, int out_dtype, int64_t ldo, double* stats, int32_t B, int32_t H,
                          int32_t W, int32_t C, int flip, const float* addend, int64_t ldadd, cudaStream_t st) {
  SBM_CHECK_ARG(x && w && out && B > 0 && C > 0 && H > 0 && W > 0, "sbm_dwconv7: bad args");
  // bf16 output on 16x16 / 8x8 / 4x4 maps: banded products on the tensor cores (dwconv_mma.cu)
  if (out_dtype == SBM_BF16 && !flip && addend == nullptr) {
    const int rc = dwconv7_mma_launch(x, ldx, w, bias, cond, ldc, out, ldo, stats, B, H, W, C, st);
    if (rc >= 0) return rc;
  }
  if (H == 1 && W == 1 && addend == nullptr) {
    const int blocks = (B + 7) / 8;
    if (out_dtype == SBM_BF16)
      dwconv7_point_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(x, ldx, w, bias, cond, ldc, (__nv_bfloat16*)out, ldo, stats, B, C);
    else
      dwconv7_point_kernel<float><<<blocks, 256, 0, st>>>(x, ldx, w, bias, cond, ldc, (float*)out, ldo, stats, B, C);
    SBM_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
  }
  const bool aligned = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (aligned && H == W && (W == 1 || W == 2 || W == 4 || W == 8 || W == 16)) {
#define SBM_DW_PIPE(WW)                                                                                              \
  return out_dtype == SBM_BF16                                                                                       \
             ? dwconv7_pipe_launch<WW, __nv_bfloat16>(x, ldx, w, bias, cond, ldc, out, ldo, stats, B, H, C, flip,   \
                                                      addend, ldadd, st)                                             \
             : dwconv7_pipe_launch<WW, float>(x, ldx, w, bias, cond, ldc, out, ldo, stats, B, H, C, flip, addend,   \
                                              ldadd, st)
    switch (W) {
      case 1: SBM_DW_PIPE(1);
      case 2: SBM_DW_PIPE(2);
      case 4: SBM_DW_PIPE(4);
      case 8: SBM_DW_PIPE(8);
      default: SBM_DW_PIPE(16);
    }
#undef SBM_DW_PIPE
  }
  SBM_CHECK_ARG(out_dtype == SBM_F32, "sbm_dwconv7: bf16 output needs a square power-of-two map <= 16");
  float* outf = (float*)out;
  const size_t smem = ((size_t)H * W * 33 + 49 * kDwCh) * sizeof(float);
  SBM_CHECK_ARG(smem <= 200 * 1024, "sbm_dwconv7: %dx%d map does not fit the shared-memory slab", H, W);
  dim3 grid((C + kDwCh - 1) / kDwCh, B);
  const size_t smem_rows = (size_t)H * W * 33 * sizeof(float);
  const int threads = 32 * std::max(1, std::min(8, H));
#define SBM_DW_ROWS(WW)                                                                                             \
  dwconv7_rows_kernel<WW><<<grid, threads, smem_rows, st>>>(x, ldx, w, bias, cond, ldc, outf, ldo, stats, C, H, flip, \
                                                            addend, ldadd)
  if (W == 16 && smem_rows <= 48 * 1024) SBM_DW_ROWS(16);
  else if (W == 8) SBM_DW_ROWS(8);
  else if (W == 4) SBM_DW_ROWS(4);
  else if (W == 2) SBM_DW_ROWS(2);
  else if (W == 1) SBM_DW_ROWS(1);
  else {
    SBM_CHECK_ARG(!flip && !addend, "sbm_dwconv7_bwd_input: unsupported width %d", W);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      SBM_CUDA_OK(cudaFuncSetAttribute(dwconv7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    dwconv7_kernel<<<grid, 256, smem, st>>>(x, ldx, w, bias, cond, ldc, outf, ldo, stats, C, H, W);
  }
#undef SBM_DW_ROWS
  SBM_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

template <typename TIn>
static int linear_attn_mma_launch(const TIn* qkv, int64_t ldq, void* out, int64_t ldo, int B, int n, int heads, float scale,
                                  cudaStream_t st) {
  const size_t sm = (size_t)3 * n * 64 + 32 * 64 + 2 * 8 * 32 * 4 + 32 * 4;
  static bool configured = false;
  if (!configured) {
    SBM_CUDA_OK(cudaFuncSetAttribute(linear_attn_mma_kernel<8, TIn>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     3 * 256 * 64 + 32 * 64 + 2 * 8 * 32 * 4 + 32 * 4));
    configured = true;
  }
  dim3 grid(heads, B);
  if (n == 256)
    linear_attn_mma_kernel<8, TIn><<<grid, 256, sm, st>>>(qkv, ldq, (__nv_bfloat16*)out, ldo, heads, scale);
  else
    linear_attn_mma_kernel<2, TIn><<<grid, 256, sm, st>>>(qkv, ldq, (__nv_bfloat16*)out, ldo, heads, scale);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace sbm

using namespace sbm;

extern "C" {

int sbm_stem_im2col(const float* x, void* a, int32_t B, int32_t C, int32_t H, int32_t W, int32_t kh, int32_t kw,
                    int32_t ldk, void* stream) {
  SBM_CHECK_ARG(x && a && B > 0 && C > 0 && ldk >= C * kh * kw, "sbm_stem_im2col: bad args");
  SBM_CHECK_ARG(ldk % 8 == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0, "sbm_stem_im2col: ldk must be a multiple of 8");
  const int64_t total = (int64_t)B * H * W * (ldk / 8);
  // padded image (floats, rounded up to 16 bytes) + offset table
  const size_t img = (((size_t)C * (H + kh - 1) * (W + kw - 1) + 3) & ~size_t(3)) * sizeof(float);
  const size_t smem = img + (size_t)ldk * sizeof(int);
  if (smem <= 48 * 1024 && img == (size_t)C * (H + kh - 1) * (W + kw - 1) * sizeof(float)) {
    stem_im2col_smem_kernel<<<std::min(B, sm_count() * 4), 256, smem, (cudaStream_t)stream>>>(
        x, (__nv_bfloat16*)a, B, C, H, W, kh, kw, ldk);
  } else {
    stem_im2col_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)a, B, C, H, W, kh, kw,
                                                                               ldk);
  }
  SBM_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int sbm_dwconv7_fwd(const float* x, int64_t ldx, const float* w, const float* bias, const float* cond, int64_t ldc,
                    float* out, int64_t ldo, double* stats, int32_t B, int32_t H, int32_t W, int32_t C,
                    void* stream) {
  return dwconv7_launch(x, ldx, w, bias, cond, ldc, out, SBM_F32, ldo, stats, B, H, W, C, 0, nullptr, 0,
                        (cudaStream_t)stream);
}

int sbm_dwconv7_fwd_bf16(const float* x, int64_t ldx, const float* w, const float* bias, const float* cond,
                         int64_t ldc, void* out_bf16, int64_t ldo, double* stats, int32_t B, int32_t H, int32_t W,
                         int32_t C, void* stream) {
  return dwconv7_launch(x, ldx, w, bias, cond, ldc, out_bf16, SBM_BF16, ldo, stats, B, H, W, C, 0, nullptr, 0,
                        (cudaStream_t)stream);
}

int sbm_dwconv7_bwd_input(const float* dy, int64_t lddy, const float* w, const float* addend, int64_t ldadd,
                          float* out, int64_t ldo, int32_t B, int32_t H, int32_t W, int32_t C, void* stream) {
  return dwconv7_launch(dy, lddy, w, nullptr, nullptr, 0, out, SBM_F32, ldo, nullptr, B, H, W, C, 1, addend, ldadd,
                        (cudaStream_t)stream);
}

int sbm_group_stats(const void* x, int32_t in_dtype, int64_t ldx, int32_t B, int32_t HW, int32_t C, int32_t G,
                    double* stats, void* stream) {
  SBM_CHECK_ARG(x && stats && B > 0 && G > 0 && C % G == 0, "sbm_group_stats: bad args");
  // row-wise kernel: 4-channel vectors inside one group, a pixel row fits the block, groups fit its reduction buffer
  const int cpg = C / G, tpr = C / 4;
  const int esz = in_dtype == SBM_F32 ? 4 : 2;
  static const bool rows_env = [] { const char* e = getenv("SBM_GROUP_STATS_ROWS"); return e ? atoi(e) != 0 : true; }();
  if (rows_env && C % 4 == 0 && cpg % 4 == 0 && (cpg / 4 & (cpg / 4 - 1)) == 0 && cpg / 4 <= 32 && tpr <= 256 &&
      256 % tpr == 0 && 256 / tpr <= 8 && G <= 64 && ldx % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(x) & (4 * esz - 1)) == 0) {
    // enough blocks for two waves when the batch is small, at least 32 pixels per block otherwise
    int ppb = std::max(32, (int)(((int64_t)B * HW + 2 * sm_count() * 4 - 1) / (2 * sm_count() * 4)));
    ppb = std::min(ppb, HW);
    dim3 grid2((HW + ppb - 1) / ppb, B);
    if (in_dtype == SBM_F32)
      group_stats_rows_kernel<float><<<grid2, 256, 0, (cudaStream_t)stream>>>((const float*)x, ldx, HW, C, G, ppb, stats);
    else
      group_stats_rows_kernel<__nv_bfloat16><<<grid2, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ldx, HW, C, G,
                                                                                    ppb, stats);
    SBM_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
  }
  const int64_t n = (int64_t)HW * (C / G);
  int chunks = (int)std::min<int64_t>((n + 2047) / 2048, 64);
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, G, B);
  group_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, in_dtype, ldx, HW, C, G, stats);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int sbm_groupnorm_apply(const void* x, int32_t in_dtype, int64_t ldx, const double* stats, const float* gamma,
                        const float* beta, const float* residual, int64_t ldr, void* out, int32_t out_dtype,
                        int64_t ldo, float* out_f32, int64_t ldo_f32, int32_t B, int32_t HW, int32_t C, int32_t G,
                        float eps, int32_t act, void* stream) {
  return sbm_groupnorm_apply_mod(x, in_dtype, ldx, stats, gamma, beta, residual, ldr, out, out_dtype, ldo, out_f32,
                                 ldo_f32, B, HW, C, G, eps, act, nullptr, nullptr, 0, nullptr, 0, stream);
}

int sbm_groupnorm_apply_mod(const void* x, int32_t in_dtype, int64_t ldx, const double* stats, const float* gamma,
                            const float* beta, const float* residual, int64_t ldr, void* out, int32_t out_dtype,
                            int64_t ldo, float* out_f32, int64_t ldo_f32, int32_t B, int32_t HW, int32_t C, int32_t G,
                            float eps, int32_t act, const float* mod_scale, const float* mod_shift, int64_t ld_mod,
                            const float* post_add, int64_t ld_post, void* stream) {
  SBM_CHECK_ARG(x && stats && gamma && beta && (out || out_f32) && B > 0 && G > 0 && C % G == 0,
                "sbm_groupnorm_apply: bad args");
  SBM_CHECK_ARG((mod_scale == nullptr) == (mod_shift == nullptr), "sbm_groupnorm_apply_mod: scale and shift come together");
  SBM_CHECK_ARG((mod_scale == nullptr || ld_mod >= C) && (post_add == nullptr || ld_post >= C),
                "sbm_groupnorm_apply_mod: per-sample row stride < C");
  SBM_CHECK_ARG(G <= 64, "sbm_groupnorm_apply: at most 64 groups");
  const int isz = in_dtype == SBM_F32 ? 4 : 2, osz = out_dtype == SBM_F32 ? 4 : 2;
  auto al16 = [](const void* p, int64_t ld, int esz) {
    return p == nullptr || (((reinterpret_cast<uintptr_t>(p) & 15) == 0) && ((ld * esz) % 16 == 0));
  };
  const int vec_ok = al16(x, ldx, isz) && al16(out, ldo, osz) && al16(residual, ldr, 4) && al16(out_f32, ldo_f32, 4);
  // block = (channel octets) x (pixel lanes); enough blocks to fill the machine ~8x, at most one pixel per lane trip
  const int co = (C + 7) / 8;
  const int lanes = std::max(1, 256 / std::min(co, 256));
  int chunks = (int)std::min<int64_t>((HW + lanes - 1) / lanes, std::max<int64_t>(1, (int64_t)sm_count() * 8 / B));
  if (chunks < 1) chunks = 1;
  // samples per block: small maps at large batch (one chunk per sample, only a few pixels per thread) are walked SB
  // samples at a time -- at least ~8 pixels per thread, the machine still covered 4 x, SB * G statistics slots <= 256
  int SB = 1;
  if (chunks == 1) {
    const int64_t per_sample = (int64_t)HW * std::min(co, 256);
    SB = (int)std::min<int64_t>((8 * 256 + per_sample - 1) / per_sample, 16);
    SB = std::min(SB, std::max(1, B / (sm_count() * 4)));
    SB = std::max(1, std::min(SB, 256 / G));
  }
  dim3 grid(chunks, (B + SB - 1) / SB);
  cudaStream_t s = (cudaStream_t)stream;
#define SBM_GN_ARGS(TI, TO)                                                                                       \
  (const TI*)x, ldx, stats, gamma, beta, residual, ldr, (TO*)out, ldo, out_f32, ldo_f32, HW, C, G, eps, act, vec_ok,    \
      mod_scale, mod_shift, ld_mod, post_add, ld_post, B, SB
#define SBM_GN_LAUNCH(TI, TO)                                                                                     \
  if (SB > 1 && mod) groupnorm_apply_kernel<TI, TO, true, true><<<grid, 256, 0, s>>>(SBM_GN_ARGS(TI, TO));            \
  else if (SB > 1) groupnorm_apply_kernel<TI, TO, true, false><<<grid, 256, 0, s>>>(SBM_GN_ARGS(TI, TO));             \
  else if (mod) groupnorm_apply_kernel<TI, TO, false, true><<<grid, 256, 0, s>>>(SBM_GN_ARGS(TI, TO));                \
  else groupnorm_apply_kernel<TI, TO, false, false><<<grid, 256, 0, s>>>(SBM_GN_ARGS(TI, TO))
  const bool mod = mod_scale != nullptr || post_add != nullptr;
  if (in_dtype == SBM_F32 && out_dtype == SBM_BF16) { SBM_GN_LAUNCH(float, __nv_bfloat16); }
  else if (in_dtype == SBM_F32 && out_dtype == SBM_F32) { SBM_GN_LAUNCH(float, float); }
  else if (in_dtype == SBM_BF16 && out_dtype == SBM_BF16) { SBM_GN_LAUNCH(__nv_bfloat16, __nv_bfloat16); }
  else { SBM_GN_LAUNCH(__nv_bfloat16, float); }
#undef SBM_GN_LAUNCH
#undef SBM_GN_ARGS
  SBM_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int sbm_upsample_nearest2x(const void* x, int64_t ldx, void* out, int64_t ldo, int32_t B, int32_t H, int32_t W,
                           int32_t C, void* stream) {
  SBM_CHECK_ARG(x && out && B > 0 && C > 0 && ldx % 8 == 0 && ldo % 8 == 0, "sbm_upsample_nearest2x: bad args");
  const int C8 = (C + 7) / 8;
  upsample2x_kernel<<<grid_for((int64_t)B * 4 * H * W * C8, 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)out, ldo, B, H, W, C8);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int sbm_time_embed(const float* t, void* out_bf16, float* out_f32, int32_t B, int32_t dim, int32_t ld, int32_t mode,
                   void* stream) {
  SBM_CHECK_ARG(t && out_bf16 && B > 0 && dim >= 4 && ld >= dim, "sbm_time_embed: bad args");
  time_embed_kernel<<<grid_for((int64_t)B * ld, 256), 256, 0, (cudaStream_t)stream>>>(
      t, (__nv_bfloat16*)out_bf16, out_f32, B, dim, ld, mode);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int sbm_linear_attn_fwd(const float* qkv, int64_t ldq, void* out, int64_t ldo, int32_t B, int32_t n, int32_t heads,
                        float scale, void* stream) {
  SBM_CHECK_ARG(qkv && out && B > 0 && n > 0 && heads > 0, "sbm_linear_attn_fwd: bad args");
  dim3 grid(heads, B);
  static const int use_mma = [] { const char* e = getenv("SBM_ATTN_MMA"); return e ? atoi(e) : 1; }();
  if (use_mma && (n == 256 || n == 64) && ldq % 4 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
      ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    return linear_attn_mma_launch<float>(qkv, ldq, out, ldo, B, n, heads, scale, (cudaStream_t)stream);
  }
  if (use_mma && n <= 16) {
    const int pairs = heads * B;
    if (n <= 4)
      linear_attn_small_kernel<4><<<(pairs + 7) / 8, 256, 0, (cudaStream_t)stream>>>(qkv, ldq, (__nv_bfloat16*)out, ldo, n,
                                                                                    heads, pairs, scale);
    else
      linear_attn_small_kernel<16><<<(pairs + 7) / 8, 256, 0, (cudaStream_t)stream>>>(qkv, ldq, (__nv_bfloat16*)out, ldo,
                                                                                     n, heads, pairs, scale);
    SBM_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
  }
  const int n4 = (n + 3) & ~3;
  const size_t smem_t = ((size_t)n4 * 32 + std::max(n4 * 32, 4096) + (size_t)n4 * 32 + 1024 + 512) * sizeof(float);
  if (ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0 && ldq % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && smem_t <= 110 * 1024) {
    static size_t configured = 0;
    if (smem_t > 48 * 1024 && smem_t > configured) {
      SBM_CUDA_OK(cudaFuncSetAttribute(linear_attn_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem_t));
      configured = smem_t;
    }
    linear_attn_tiled_kernel<<<grid, 256, smem_t, (cudaStream_t)stream>>>(qkv, ldq, (__nv_bfloat16*)out, ldo, n, heads,
                                                                          scale);
    SBM_CUDA_OK(cudaGetLastError());
    count_launch();
    return 0;
  }
  const size_t smem = ((size_t)3 * n * 33 + 32 * 33) * sizeof(float);
  SBM_CHECK_ARG(smem <= 200 * 1024, "sbm_linear_attn_fwd: n=%d too large", n);
  static size_t configured = 0, configured_occ = 0;
  if (smem > 48 * 1024 && smem > configured) {
    SBM_CUDA_OK(cudaFuncSetAttribute(linear_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  linear_attn_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(qkv, ldq, (__nv_bfloat16*)out, ldo, n, heads, scale);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

int sbm_linear_attn_fwd_bf16(const void* qkv, int64_t ldq, void* out, int64_t ldo, int32_t B, int32_t n, int32_t heads,
                             float scale, void* stream) {
  SBM_CHECK_ARG(qkv && out && B > 0 && heads > 0, "sbm_linear_attn_fwd_bf16: bad args");
  SBM_CHECK_ARG(n == 256 || n == 64, "sbm_linear_attn_fwd_bf16: n = %d (the bf16-input kernel covers n = 64 and 256)", n);
  SBM_CHECK_ARG(ldq % 4 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 7) == 0 && ldo % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(out) & 15) == 0, "sbm_linear_attn_fwd_bf16: misaligned operands");
  return linear_attn_mma_launch<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(qkv), ldq, out, ldo, B, n, heads, scale,
                                               (cudaStream_t)stream);
}

int sbm_softmax_attn_fwd(const float* qkv, int64_t ldq, void* out, int64_t ldo, int32_t B, int32_t n, int32_t heads,
                         int32_t dh, int32_t q_off, int32_t k_off, int32_t v_off, int32_t head_stride, float scale,
                         void* stream) {
  SBM_CHECK_ARG(qkv && out && B > 0 && n > 0 && heads > 0 && dh > 0, "sbm_softmax_attn_fwd: bad args");
  const size_t smem = ((size_t)n * (n + 1) + 2 * (size_t)n * 33) * sizeof(float);
  SBM_CHECK_ARG(smem <= 200 * 1024, "sbm_softmax_attn_fwd: n=%d too large for the shared-memory score tile", n);
  static size_t configured = 0, configured_occ = 0;
  if (smem > 48 * 1024 && smem > configured) {
    SBM_CUDA_OK(cudaFuncSetAttribute(softmax_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid(heads, B);
  softmax_attn_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(qkv, ldq, (__nv_bfloat16*)out, ldo, n, dh, q_off,
                                                                 k_off, v_off, head_stride, scale);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch();
  return 0;
}

}  // extern "C"
