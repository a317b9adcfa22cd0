// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM).
//
//   out[b, oh, ow, n] = epilogue( sum_{tap, c} X[b, ih(tap), iw(tap), c] * Wpk[tap][n][c] )
//
// GEMM view: M = batch*OH*OW output pixels (128 per CTA), N = cout (BN per CTA),
// K = taps * cin walked in 64-channel blocks.  There is no im2col buffer: for every
// (tap, channel block) one TMA box load fetches the 128 shifted input pixels x 64 channels
// straight from the channels-last activation tensor; pixels that fall in the zero padding
// are outside the tensor-map extent and TMA fills them with zeros.  Taps that only ever see
// padding (e.g. 8 of the 9 taps of a 3x3 convolution on a 1x1 map) are dropped from the K loop.
//
// One generic 5-D activation view serves every layer type:
//     (c, wv, q, hv, b)  with element strides (1, sw, sq, sh, sb)
//   stride-1 conv / linear : wv=w, hv=h, q unused
//   stride-2 conv          : wv=w/2, hv=h/2, q = (h parity)*W + (w parity)   (space-to-depth by strides)
//   ConvTranspose 4x4 s2   : four output-parity phases (grid.z), each a 2x2-tap stride-1 conv
//
// Warp roles (256 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM -> registers -> bias/act/residual -> global).
//
// Replaces: nn.Conv2d / nn.ConvTranspose2d / nn.Linear in unet_model.py:30,33,107,110,113,132-133,
// 157-159,208,224,226,272 and unet_openai.py:185,207,253-268,322-324,421-425.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "../../include/sbmae_b200.h"
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace sbm {

std::atomic<unsigned long long> g_launches{0};
static bool g_force_single = false;  // debugging / A-B timing switch (sbm_conv_force_single_cta)
static thread_local int g_last_variant = 0;  // BN | pair << 16 | staged << 17 of the last launch (bench bookkeeping)
static int g_pixel_major = [] { const char* e = getenv("SBM_PIXEL_MAJOR"); return e ? atoi(e) : -1; }();       // -1: by work estimate, 0: never, 1: whenever the CTA-pair kernel runs the layer
static bool g_force_direct = false;  // A-B switch: per-thread global stores instead of the TMA-staged epilogue

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kMaxTaps = 16;

struct TapTable {
  int32_t ntaps;
  int32_t out_off;          // element offset of this phase inside the output tensor
  int32_t out2_off;         // same for the optional bf16 copy
  int32_t out_q;            // coordinate of this phase in the parity dimension of the output tensor maps
  int8_t dh[kMaxTaps];      // added to the tile's hv origin
  int8_t dw[kMaxTaps];      // wv coordinate of the box start
  int16_t q[kMaxTaps];      // coordinate in the parity dimension
  int16_t wtap[kMaxTaps];   // tap index inside the packed weight tensor
};

struct ConvKernelParams {
  int32_t batch;
  int32_t log_ow, log_th;   // tile = nb images x 2^log_th rows x 2^log_ow columns = 128 pixels
  int32_t log_oh;
  int32_t cin, cout;
  int32_t cblocks;          // ceil(cin / 64)
  int64_t o_sb, o_sh, o_sw, o_sc;  // output element strides (batch, row, col, channel)
  int64_t r_sb, r_sh, r_sw;        // residual element strides
  int64_t o2_sb, o2_sh, o2_sw;     // optional second (bf16) output
  const float* bias;
  const void* residual;
  void* out;
  void* out2;
  double* stats;
  int32_t act, out_dtype, res_dtype, vec_ok;
  int32_t out2_preact, bias_vec;
  const float* rowbias;     // optional per-sample bias [batch][ld_rowbias] added before the activation
  int64_t ld_rowbias;
  // GroupNorm(1, cin) of the INPUT folded into the convolution (weights carry gamma; see sbm_conv_fold_groupnorm):
  //   y = rstd_b * (acc - mean_b * Sg[cls][n]) + Tb[cls][n],  cls = which 3x3 taps see real pixels at this position
  const double* gn_stats;   // [batch][2] (sum, sum of squares) of the input tensor, or NULL
  const float* gn_tab;      // [2][16][cout]: Sg then Tb
  double gn_inv_count;
  float gn_eps;
  // pixel-major tiling (stride-1 'same' convolutions at large batch): the 128 rows of a tile are 128 SAMPLES at ONE
  // output pixel, so the taps that read zero padding at that pixel are skipped for the whole tile
  int32_t pm;
  int32_t pm_blocks;        // 256-sample blocks in the batch
  int32_t pm_global;        // tile order: 1 = (pixel rank, block) -- cost-sorted over the whole list, for short lists;
                            // 0 = (block, pixel rank) -- a block's pixels stay together (DRAM page / L2 locality)
  uint8_t pm_pix[256];      // output pixels (i * W + j) ordered by falling tap count: the tile list is cost-sorted
  TapTable taps[4];
};

// per-thread (= per output row) constants of the folded GroupNorm
struct GnRow {
  float mu, rstd;
  const float* sg;
  const float* tb;
};
__device__ __forceinline__ GnRow gn_row(const ConvKernelParams& p, int b, int i, int j, bool row_ok) {
  GnRow g;
  g.mu = 0.f; g.rstd = 1.f; g.sg = p.gn_tab; g.tb = p.gn_tab;
  if (p.gn_tab == nullptr) return g;
  const int H = 1 << p.log_oh, W = 1 << p.log_ow;
  int cls = 0;
  if (row_ok) {
    const double s1 = p.gn_stats[2 * (int64_t)b], s2 = p.gn_stats[2 * (int64_t)b + 1];
    const double mean = s1 * p.gn_inv_count;
    const double var = fmax(s2 * p.gn_inv_count - mean * mean, 0.0);
    g.mu = (float)mean;
    g.rstd = rsqrtf((float)var + p.gn_eps);  // fp32 like torch's GroupNorm; the fp64 divide + sqrt was 11 % of the epilogue
    cls = (i >= 1 ? 1 : 0) | (i <= H - 2 ? 2 : 0) | (j >= 1 ? 4 : 0) | (j <= W - 2 ? 8 : 0);
  }
  g.sg = p.gn_tab + (int64_t)cls * p.cout;
  g.tb = p.gn_tab + (int64_t)(16 + cls) * p.cout;
  return g;
}
__device__ __forceinline__ void gn_apply16(const ConvKernelParams& p, const GnRow& g, float* f, int n) {
  if (p.gn_tab == nullptr) return;
  const int cmax = p.cout - 1;
  if (n + 16 <= p.cout && (p.cout & 3) == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(g.sg + n) + k);
      const float4 t = __ldg(reinterpret_cast<const float4*>(g.tb + n) + k);
      f[4 * k] = fmaf(g.rstd, f[4 * k] - g.mu * a.x, t.x);
      f[4 * k + 1] = fmaf(g.rstd, f[4 * k + 1] - g.mu * a.y, t.y);
      f[4 * k + 2] = fmaf(g.rstd, f[4 * k + 2] - g.mu * a.z, t.z);
      f[4 * k + 3] = fmaf(g.rstd, f[4 * k + 3] - g.mu * a.w, t.w);
    }
  } else {
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int c = min(n + e, cmax);
      f[e] = fmaf(g.rstd, f[e] - g.mu * __ldg(g.sg + c), __ldg(g.tb + c));
    }
  }
}

// bias + activation + residual + (bf16 rounding) + statistics + stores for 16 consecutive output channels of one
// output pixel (row); `v` holds the fp32 accumulators read from TMEM.
__device__ __forceinline__ void epilogue16(const ConvKernelParams& p, const uint32_t* v, int n, bool row_ok, int b,
                                           int64_t o_base, int64_t r_base, int64_t o2_base, float& s1, float& s2,
                                           const GnRow& gr) {
  if (n >= p.cout) return;  // warp-uniform
  float f[16];
  const bool full = (n + 16 <= p.cout);
#pragma unroll
  for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]);
  gn_apply16(p, gr, f, n);
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    float x = f[e];
    if (p.bias != nullptr && (full || n + e < p.cout)) x += __ldg(p.bias + n + e);
    if (p.rowbias != nullptr && row_ok && (full || n + e < p.cout)) x += __ldg(p.rowbias + (int64_t)b * p.ld_rowbias + n + e);
    if (p.out2_preact && row_ok && (full || n + e < p.cout))
      reinterpret_cast<__nv_bfloat16*>(p.out2)[o2_base + n + e] = __float2bfloat16_rn(x);
    if (p.act == SBM_ACT_GELU) x = gelu_exact(x);
    else if (p.act == SBM_ACT_SILU) x = silu(x);
    f[e] = x;
  }
  if (row_ok) {
    if (p.residual != nullptr) {
      if (p.res_dtype == SBM_F32) {
        const float* rp = reinterpret_cast<const float*>(p.residual) + r_base + n;
        if (full && p.vec_ok) {
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            const float4 t = *reinterpret_cast<const float4*>(rp + e);
            f[e] += t.x; f[e + 1] += t.y; f[e + 2] += t.z; f[e + 3] += t.w;
          }
        } else {
#pragma unroll
          for (int e = 0; e < 16; ++e)
            if (n + e < p.cout) f[e] += rp[e];
        }
      } else {
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + r_base + n;
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (n + e < p.cout) f[e] += __bfloat162float(rp[e]);
      }
    }
    if (p.out_dtype == SBM_BF16) {
#pragma unroll
      for (int e = 0; e < 16; ++e) f[e] = __bfloat162float(__float2bfloat16_rn(f[e]));
    }
    if (p.stats != nullptr) {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (full || n + e < p.cout) { s1 += f[e]; s2 += f[e] * f[e]; }
    }
    if (p.out_dtype == SBM_F32) {
      float* op = reinterpret_cast<float*>(p.out) + o_base;
      if (full && p.vec_ok && p.o_sc == 1) {
#pragma unroll
        for (int e = 0; e < 16; e += 4)
          *reinterpret_cast<float4*>(op + n + e) = make_float4(f[e], f[e + 1], f[e + 2], f[e + 3]);
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (n + e < p.cout) op[(int64_t)(n + e) * p.o_sc] = f[e];
      }
    } else {
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + o_base + n;
      if (full && p.vec_ok) {
        uint32_t w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
          w[e] = *reinterpret_cast<uint32_t*>(&t);
        }
        *reinterpret_cast<uint4*>(op) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(op + 8) = make_uint4(w[4], w[5], w[6], w[7]);
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (n + e < p.cout) op[e] = __float2bfloat16_rn(f[e]);
      }
    }
    if (p.out2 != nullptr && !p.out2_preact) {
      __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out2) + o2_base + n;
      if (full && p.vec_ok) {
        uint32_t w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
          w[e] = *reinterpret_cast<uint32_t*>(&t);
        }
        *reinterpret_cast<uint4*>(op) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(op + 8) = make_uint4(w[4], w[5], w[6], w[7]);
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (n + e < p.cout) op[e] = __float2bfloat16_rn(f[e]);
      }
    }
  }
}

// ---- staged epilogue (CTA-pair kernel): the thread's 16 output values go to a swizzled shared-memory row so that
// global memory only ever sees TMA box transfers (full 32-byte sectors, no per-thread strided stores).
//   fp32 rows: 64 B, CU_TENSOR_MAP_SWIZZLE_64B : 16-byte chunk k of row r lives at chunk k ^ ((r >> 1) & 3)
//   bf16 rows: 32 B, CU_TENSOR_MAP_SWIZZLE_32B : 16-byte chunk k of row r lives at chunk k ^ ((r >> 2) & 1)
// (buffers are 1024-byte aligned, so the swizzle's address bits are the row bits above).
__device__ __forceinline__ float4* stg_f32(uint8_t* buf, int r, int k) {
  return reinterpret_cast<float4*>(buf + r * 64 + ((k ^ ((r >> 1) & 3)) << 4));
}
__device__ __forceinline__ uint4* stg_bf16(uint8_t* buf, int r, int k) {
  return reinterpret_cast<uint4*>(buf + r * 32 + ((k ^ ((r >> 2) & 1)) << 4));
}
__device__ __forceinline__ void stg_store_bf16_row(uint8_t* buf, int r, const float* f) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat162 t = __floats2bfloat162_rn(f[8 * k + 2 * e], f[8 * k + 2 * e + 1]);
      w[e] = *reinterpret_cast<uint32_t*>(&t);
    }
    *stg_bf16(buf, r, k) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
// `stage` holds the residual chunk on entry (when has_res) and the output chunk on exit; `stage2` receives the bf16 copy.
__device__ __forceinline__ void epilogue_chunk_staged(const ConvKernelParams& p, const uint32_t* v, int n, bool row_ok,
                                                      int b, int lane, uint8_t* stage, uint8_t* stage2, bool has_res,
                                                      float& s1, float& s2, const GnRow& gr) {
  float f[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]);
  gn_apply16(p, gr, f, n);
  const int cmax = p.cout - 1;
  if (p.bias != nullptr) {
    if (p.bias_vec && n + 16 <= p.cout) {  // 4 broadcast 16-byte loads instead of 16 clamped scalar ones
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + n) + k);
        f[4 * k] += t.x; f[4 * k + 1] += t.y; f[4 * k + 2] += t.z; f[4 * k + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) f[e] += __ldg(p.bias + min(n + e, cmax));
    }
  }
  if (p.rowbias != nullptr && row_ok) {
    const float* rb = p.rowbias + (int64_t)b * p.ld_rowbias;
#pragma unroll
    for (int e = 0; e < 16; ++e) f[e] += __ldg(rb + min(n + e, cmax));
  }
  if (p.out2_preact) stg_store_bf16_row(stage2, lane, f);
  if (p.act == SBM_ACT_GELU) {
#pragma unroll
    for (int e = 0; e < 16; ++e) f[e] = gelu_exact(f[e]);
  } else if (p.act == SBM_ACT_SILU) {
#pragma unroll
    for (int e = 0; e < 16; ++e) f[e] = silu(f[e]);
  }
  if (has_res) {
    if (p.res_dtype == SBM_F32) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 t = *stg_f32(stage, lane, k);
        f[4 * k] += t.x; f[4 * k + 1] += t.y; f[4 * k + 2] += t.z; f[4 * k + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const uint4 t = *stg_bf16(stage, lane, k);
        const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&u[e]);
          f[8 * k + 2 * e] += __low2float(h);
          f[8 * k + 2 * e + 1] += __high2float(h);
        }
      }
    }
    __syncwarp();  // every lane has read its residual row before any lane overwrites the buffer
  }
  if (p.out_dtype == SBM_BF16) {
#pragma unroll
    for (int e = 0; e < 16; ++e) f[e] = __bfloat162float(__float2bfloat16_rn(f[e]));
  }
  if (p.stats != nullptr && row_ok) {
    if (n + 16 <= p.cout) {
#pragma unroll
      for (int e = 0; e < 16; ++e) { s1 += f[e]; s2 = fmaf(f[e], f[e], s2); }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (n + e <= cmax) { s1 += f[e]; s2 += f[e] * f[e]; }
    }
  }
  if (p.out_dtype == SBM_F32) {
#pragma unroll
    for (int k = 0; k < 4; ++k) *stg_f32(stage, lane, k) = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
  } else {
    stg_store_bf16_row(stage, lane, f);
  }
  if (p.out2 != nullptr && !p.out2_preact) stg_store_bf16_row(stage2, lane, f);
}

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = STAGES * kStageBytes;
  static constexpr int kTotal = kBarOffset + (2 * STAGES + 1) * 8 + 16 + 1024;  // + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(256, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ ConvKernelParams p) {
  using L = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const TapTable& tt = p.taps[blockIdx.z];

  // ---- tile origin
  const int log_ohw = p.log_oh + p.log_ow;
  int b0, oh0;
  if (log_ohw >= 7) {
    const int tiles_per_img = 1 << (log_ohw - 7);
    b0 = blockIdx.x / tiles_per_img;
    oh0 = (blockIdx.x % tiles_per_img) << p.log_th;
  } else {
    b0 = blockIdx.x << (7 - log_ohw);
    oh0 = 0;
  }
  const int n0 = blockIdx.y * BN;
  const int num_kb = tt.ntaps * p.cblocks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<(BN < 32 ? 32 : BN)>(tmem_slot);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int tap = kb / p.cblocks;
        const int cb = kb - tap * p.cblocks;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * L::kStageBytes;
        uint8_t* sb = sa + L::kABytes;
        ptx::mbar_expect_tx(&full_bar[stage], L::kStageBytes);
        ptx::tma_load_5d(sa, &tmA, &full_bar[stage], cb * kBK, tt.dw[tap], tt.q[tap], oh0 + tt.dh[tap], b0);
        ptx::tma_load_3d(sb, &tmB, &full_bar[stage], cb * kBK, n0, tt.wtap[tap]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after_sync();
        const uint32_t sa = ptx::smem_u32(smem + stage * L::kStageBytes);
        const uint64_t adesc = ptx::make_desc_k_sw128(sa);
        const uint64_t bdesc = ptx::make_desc_k_sw128(sa + L::kABytes);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) {
          // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in the 16-byte address field
          ptx::umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      ptx::umma_commit(tmem_full_bar);
    }
  }
  // ===================== epilogue: ALL 8 warps (the producer / issuer warps join once their loops are done; a small
  // problem runs one tile per CTA, so the epilogue is pure latency).  One output pixel (row) per thread; warp w reads
  // TMEM lane quarter w % 4 and the column half w / 4.
  __syncwarp();
  {
    const int ew = warp & 3;
    const int hc = warp >> 2;
    const int r = ew * 32 + lane;
    const int ow_mask = (1 << p.log_ow) - 1;
    const int j = r & ow_mask;
    const int i = (r >> p.log_ow) & ((1 << p.log_th) - 1);
    const int bl = r >> (p.log_ow + p.log_th);
    const int b = b0 + bl;
    const int oh = oh0 + i;
    const bool row_ok = b < p.batch;
    const int64_t o_base = (int64_t)b * p.o_sb + (int64_t)oh * p.o_sh + (int64_t)j * p.o_sw + tt.out_off;
    const int64_t r_base = (int64_t)b * p.r_sb + (int64_t)oh * p.r_sh + (int64_t)j * p.r_sw;
    const int64_t o2_base = (int64_t)b * p.o2_sb + (int64_t)oh * p.o2_sh + (int64_t)j * p.o2_sw + tt.out2_off;
    float s1 = 0.f, s2 = 0.f;
    const GnRow gr = gn_row(p, b, oh, j, row_ok);

    ptx::mbar_wait(tmem_full_bar, 0);
    ptx::tc_fence_after_sync();
    __syncwarp();

    constexpr int kHalf = BN >= 32 ? BN / 2 : BN;  // BN = 32 -> 16 columns per half
#pragma unroll 1
    for (int c0 = hc * kHalf; c0 < (hc + 1) * kHalf && c0 < BN; c0 += 16) {
      uint32_t v[16];
      ptx::tmem_ld16(tmem_base + (uint32_t(ew * 32) << 16) + c0, v);
      ptx::tmem_ld_wait();
      epilogue16(p, v, n0 + c0, row_ok, b, o_base, r_base, o2_base, s1, s2, gr);
    }

    if (p.stats != nullptr) {
      // rows of one warp are 32 consecutive pixels: they belong to one sample when OH*OW >= 32,
      // otherwise to 32/(OH*OW) samples -> segmented butterfly over groups of OH*OW lanes.
      if (!row_ok) { s1 = 0.f; s2 = 0.f; }
      const int seg = log_ohw >= 5 ? 32 : (1 << log_ohw);
      for (int o = seg >> 1; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if ((lane & (seg - 1)) == 0 && row_ok) {
        atomicAdd(p.stats + 2 * (int64_t)b, (double)s1);
        atomicAdd(p.stats + 2 * (int64_t)b + 1, (double)s2);
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<(BN < 32 ? 32 : BN)>(tmem_base);
}


// =====================================================================================================
// CTA-pair kernel (cta_group::2): two SMs of a TPC compute one 256 x BN output tile with ONE
// tcgen05.mma per K step.  Each CTA stages its own 128 activation rows and HALF of the weight rows, so the
// L2 -> shared-memory traffic per FLOP drops by 1.5x versus the single-CTA kernel (which measured L2-bound).
// Persistent: one cluster per SM pair walks a static tile list; the accumulator is double-buffered in TMEM
// (2 x BN columns) so the epilogue of tile i overlaps the main loop of tile i+1.
//   warp 0: TMA producer (both CTAs)      warp 1: MMA issuer (leader CTA only)
//   warp 2: TMEM allocator                warps 4..11: epilogue (lane quarter = warp%4, column half = (warp-4)/4)
// =====================================================================================================
// Epilogue staging (kStaged): every epilogue warp owns 3 x 2 KB buffers (residual chunk in / output chunk out,
// 32 rows x 16 fp32 columns, 64-byte swizzle) and 2 x 1 KB buffers (bf16 copy, 32-byte swizzle).
constexpr int kEC = 16;                 // output columns per epilogue chunk (= one tcgen05.ld.32x32b.x16)
constexpr int kStgMain = 2048;
constexpr int kStgOut2 = 1024;
constexpr int kStgPerWarp = 3 * kStgMain + 2 * kStgOut2;  // 8 KB
constexpr int kEpiWarps = 8;

template <int BN, int STAGES, bool kStaged>
struct Smem2Layout {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = (BN / 2) * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagingOffset = STAGES * kStageBytes;
  static constexpr int kBarOffset = kStagingOffset + (kStaged ? kEpiWarps * kStgPerWarp : 0);
  static constexpr int kTotal = kBarOffset + (2 * STAGES + 4 + 3 * kEpiWarps) * 8 + 16 + 1024;
};

struct PairSchedule {
  int32_t m_tiles, m_pairs, n_tiles, nphase, total;
};

struct EpiMaps {
  CUtensorMap out, res, out2;  // 32-row x kEC-column boxes of the output / residual / bf16-copy tensors
};

template <int BN, int STAGES, bool kStaged>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
conv_igemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ EpiMaps em, const __grid_constant__ ConvKernelParams p,
                       const PairSchedule sch) {
  using L = Smem2Layout<BN, STAGES, kStaged>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;       // [2]
  uint64_t* res_bar = tempty_bar + 2;         // [kEpiWarps][3] residual-chunk arrival (staged epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 3 * kEpiWarps);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  const int log_ohw = p.log_oh + p.log_ow;
  const int per_phase = sch.m_pairs * sch.n_tiles;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull_bar[a], 1);
      ptx::mbar_init(&tempty_bar[a], 16);  // 8 epilogue warps x 2 CTAs arrive on the leader's barrier
    }
    for (int a = 0; a < 3 * kEpiWarps; ++a) ptx::mbar_init(&res_bar[a], 1);
    ptx::fence_mbar_init();
  }
  if (kStaged && warp == 3 && lane == 0) {
    ptx::prefetch_tmap(&em.out);
    if (p.residual != nullptr) ptx::prefetch_tmap(&em.res);
    if (p.out2 != nullptr) ptx::prefetch_tmap(&em.out2);
  }
  if (warp == 2) ptx::tmem_alloc2<2 * BN>(tmem_slot);
  ptx::tc_fence_before_sync();
  ptx::cluster_sync_all();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  // tile t -> phase, N tile, first sample, first output row (and, pixel-major: the output column pj of the tile)
  auto tile_origin = [&](int t, int& ph, int& nt, int& b0, int& oh0, int& pj) {
    ph = t / per_phase;
    const int rem = t - ph * per_phase;
    const int mp = rem / sch.n_tiles;
    nt = rem - mp * sch.n_tiles;
    pj = 0;
    if (p.pm) {  // m-pair index -> (pixel rank, 256-sample block); interior pixels (most taps) first, corners last
      int pr, sb;
      if (p.pm_global) {
        pr = mp / p.pm_blocks;
        sb = mp - pr * p.pm_blocks;
      } else {
        sb = mp >> log_ohw;
        pr = mp & ((1 << log_ohw) - 1);
      }
      const int px = p.pm_pix[pr];
      b0 = (sb << 8) + ((int)rank << 7);
      oh0 = px >> p.log_ow;
      pj = px & ((1 << p.log_ow) - 1);
      return;
    }
    const int mt = 2 * mp + (int)rank;
    if (log_ohw >= 7) {
      const int tiles_per_img = 1 << (log_ohw - 7);
      b0 = mt / tiles_per_img;
      oh0 = (mt % tiles_per_img) << p.log_th;
    } else {
      b0 = mt << (7 - log_ohw);
      oh0 = 0;
    }
  };
  // k-th tile of this cluster.  Pixel-major tile lists are sorted by falling cost and dealt out in snake order (round
  // k runs over the clusters forwards, round k+1 backwards), which evens out the per-cluster sums of unequal tiles
  auto tile_of_round = [&](int k) -> int {
    return k * npairs + ((p.pm && (k & 1)) ? npairs - 1 - pair : pair);
  };
  // taps of the table that read at least one real pixel for this tile (pixel-major: the tile is one output pixel)
  auto tap_mask = [&](const TapTable& tt, int oh0, int pj) -> uint32_t {
    if (!p.pm) return (1u << tt.ntaps) - 1u;
    uint32_t m = 0;
    for (int k = 0; k < tt.ntaps; ++k)
      if ((unsigned)(oh0 + tt.dh[k]) < (1u << p.log_oh) && (unsigned)(pj + tt.dw[k]) < (1u << p.log_ow)) m |= 1u << k;
    return m;
  };

  if (warp == 0) {
    // ===================== TMA producer (each CTA loads its own operand halves)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int k = 0; k * npairs < sch.total; ++k) {
        const int t = tile_of_round(k);
        if (t >= sch.total) continue;
        int ph, nt, b0, oh0, pj;
        tile_origin(t, ph, nt, b0, oh0, pj);
        const TapTable& tt = p.taps[ph];
        const int n0 = nt * BN + (int)rank * (BN / 2);
        for (uint32_t tm = tap_mask(tt, oh0, pj); tm != 0; tm &= tm - 1) {
          const int tap = __ffs(tm) - 1;
          for (int cb = 0; cb < p.cblocks; ++cb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * L::kStageBytes;
            uint8_t* sb = sa + L::kABytes;
            if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
            const uint32_t lead_bar = ptx::mapa_u32(ptx::smem_u32(&full_bar[stage]), 0);
            ptx::tma_load_5d_2sm(sa, &tmA, lead_bar, cb * kBK, pj + tt.dw[tap], tt.q[tap], oh0 + tt.dh[tap], b0);
            ptx::tma_load_3d_2sm(sb, &tmB, lead_bar, cb * kBK, n0, tt.wtap[tap]);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread)
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(256, BN);
      int stage = 0, astage = 0;
      uint32_t phase = 0, aphase = 0;
      for (int k = 0; k * npairs < sch.total; ++k) {
        const int t = tile_of_round(k);
        if (t >= sch.total) continue;
        int ph, nt, b0, oh0, pj;
        tile_origin(t, ph, nt, b0, oh0, pj);
        const int num_kb = __popc(tap_mask(p.taps[ph], oh0, pj)) * p.cblocks;
        ptx::mbar_wait(&tempty_bar[astage], aphase ^ 1);
        ptx::tc_fence_after_sync();
        const uint32_t tacc = tmem_base + (uint32_t)(astage * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t sa = ptx::smem_u32(smem + stage * L::kStageBytes);
          const uint64_t adesc = ptx::make_desc_k_sw128(sa);
          const uint64_t bdesc = ptx::make_desc_k_sw128(sa + L::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            ptx::umma_bf16_2cta(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit_2cta(&empty_bar[stage], 3);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit_2cta(&tfull_bar[astage], 3);
        astage ^= 1;
        if (astage == 0) aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue
    const int e = warp - 4;
    const int ew = e & 3;      // TMEM lane quarter
    const int hc = e >> 2;     // column half
    const int r = ew * 32 + lane;
    // row r of a tile = pixel (i, j) of local image bl (pixel-major: sample r at the tile's pixel)
    const int j_r = p.pm ? 0 : (r & ((1 << p.log_ow) - 1));
    const int i = p.pm ? 0 : ((r >> p.log_ow) & ((1 << p.log_th) - 1));
    const int bl = p.pm ? r : (r >> (p.log_ow + p.log_th));
    const uint32_t lead_tempty0 = ptx::mapa_u32(ptx::smem_u32(&tempty_bar[0]), 0);
    // staged path: this warp's 32 rows are one TMA box (columns, ow-run, 1, rows, images) starting at
    const int sub_j = p.pm ? 0 : ((ew * 32) & ((1 << p.log_ow) - 1));
    const int sub_i = p.pm ? 0 : (((ew * 32) >> p.log_ow) & ((1 << p.log_th) - 1));
    const int sub_b = p.pm ? ew * 32 : ((ew * 32) >> (p.log_ow + p.log_th));
    uint8_t* wst = smem + L::kStagingOffset + e * kStgPerWarp;
    uint64_t* rbar = res_bar + 3 * e;
    const bool has_res = p.residual != nullptr;
    const uint32_t res_bytes = 32u * kEC * (p.res_dtype == SBM_F32 ? 4u : 2u);
    uint32_t nchunk = 0;  // chunks this warp has staged so far (buffer rotation + barrier parity)
    int astage = 0;
    uint32_t aphase = 0;
    for (int k = 0; k * npairs < sch.total; ++k) {
        const int t = tile_of_round(k);
        if (t >= sch.total) continue;
      int ph, nt, b0, oh0, pj;
      tile_origin(t, ph, nt, b0, oh0, pj);
      const TapTable& tt = p.taps[ph];
      const int b = b0 + bl;
      const int oh = oh0 + i;
      const int j = pj + j_r;
      const bool row_ok = b < p.batch;
      float s1 = 0.f, s2 = 0.f;
      const GnRow gr = gn_row(p, b, oh, j, row_ok);
      if constexpr (kStaged) {
        const int cj = pj + sub_j, ci = oh0 + sub_i, cb = b0 + sub_b, cq = tt.out_q;
        const int ncol0 = nt * BN + hc * (BN / 2);
        const int nch = min((BN / 2) / kEC, max(0, (p.cout - ncol0 + kEC - 1) / kEC));
        if (has_res && nch > 0 && lane == 0) {
          ptx::bulk_wait_group_read<1>();
          uint64_t* rb = &rbar[nchunk % 3];
          ptx::mbar_expect_tx(rb, res_bytes);
          ptx::tma_load_5d(wst + (nchunk % 3) * kStgMain, &em.res, rb, ncol0, cj, 0, ci, cb);
        }
        ptx::mbar_wait(&tfull_bar[astage], aphase);
        ptx::tc_fence_after_sync();
        const uint32_t tacc = tmem_base + (uint32_t)(astage * BN) + (uint32_t(ew * 32) << 16) + (uint32_t)(hc * (BN / 2));
#pragma unroll 1
        for (int c = 0; c < nch; ++c) {
          const int col0 = ncol0 + c * kEC;
          uint32_t v[16];
          __syncwarp();
          ptx::tmem_ld16(tacc + c * kEC, v);
          // buffers of chunk nchunk+1 (main) / nchunk (bf16 copy) were last used by the stores of chunk nchunk-2
          if (lane == 0) {
            ptx::bulk_wait_group_read<1>();
            if (has_res && c + 1 < nch) {
              uint64_t* rb = &rbar[(nchunk + 1) % 3];
              ptx::mbar_expect_tx(rb, res_bytes);
              ptx::tma_load_5d(wst + ((nchunk + 1) % 3) * kStgMain, &em.res, rb, col0 + kEC, cj, 0, ci, cb);
            }
          }
          __syncwarp();
          uint8_t* stage = wst + (nchunk % 3) * kStgMain;
          uint8_t* stage2 = wst + 3 * kStgMain + (nchunk & 1) * kStgOut2;
          if (has_res) ptx::mbar_wait(&rbar[nchunk % 3], (nchunk / 3) & 1);
          ptx::tmem_ld_wait();
          epilogue_chunk_staged(p, v, col0, row_ok, b, lane, stage, stage2, has_res, s1, s2, gr);
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_5d(&em.out, stage, col0, cj, cq, ci, cb);
            if (p.out2 != nullptr) ptx::tma_store_5d(&em.out2, stage2, col0, cj, cq, ci, cb);
            ptx::bulk_commit_group();
          }
          ++nchunk;
        }
      } else {
        const int64_t o_base = (int64_t)b * p.o_sb + (int64_t)oh * p.o_sh + (int64_t)j * p.o_sw + tt.out_off;
        const int64_t r_base = (int64_t)b * p.r_sb + (int64_t)oh * p.r_sh + (int64_t)j * p.r_sw;
        const int64_t o2_base = (int64_t)b * p.o2_sb + (int64_t)oh * p.o2_sh + (int64_t)j * p.o2_sw + tt.out2_off;
        ptx::mbar_wait(&tfull_bar[astage], aphase);
        ptx::tc_fence_after_sync();
        const uint32_t tacc = tmem_base + (uint32_t)(astage * BN) + (uint32_t(ew * 32) << 16);
#pragma unroll 1
        for (int c0 = hc * (BN / 2); c0 < (hc + 1) * (BN / 2); c0 += 16) {
          uint32_t v0[16];
          ptx::tmem_ld16(tacc + c0, v0);
          ptx::tmem_ld_wait();
          epilogue16(p, v0, nt * BN + c0, row_ok, b, o_base, r_base, o2_base, s1, s2, gr);
        }
      }
      // accumulator stage drained: hand it back to the MMA issuer (leader CTA's barrier)
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(lead_tempty0 + (uint32_t)astage * 8u);
      if (p.stats != nullptr) {
        if (!row_ok) { s1 = 0.f; s2 = 0.f; }
        const int seg = p.pm ? 1 : (log_ohw >= 5 ? 32 : (1 << log_ohw));   // rows of this warp that share a sample
        for (int o = seg >> 1; o > 0; o >>= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if ((lane & (seg - 1)) == 0 && row_ok) {
          atomicAdd(p.stats + 2 * (int64_t)b, (double)s1);
          atomicAdd(p.stats + 2 * (int64_t)b + 1, (double)s2);
        }
      }
      astage ^= 1;
      if (astage == 0) aphase ^= 1;
    }
    // shared memory must stay valid until every bulk store has read it
    if (kStaged && lane == 0) ptx::bulk_wait_group<0>();
  }

  ptx::tc_fence_before_sync();
  ptx::cluster_sync_all();
  if (warp == 2) ptx::tmem_dealloc2<2 * BN>(tmem_base);
}

// ------------------------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

static int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return ((1 << l) == v) ? l : -1;
}

template <int BN, int STAGES>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvKernelParams& p, dim3 grid,
                       cudaStream_t stream) {
  using L = SmemLayout<BN, STAGES>;
  static bool configured = false;
  if (!configured) {
    SBM_CUDA_OK(cudaFuncSetAttribute(conv_igemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     L::kTotal));
    configured = true;
  }
  conv_igemm_kernel<BN, STAGES><<<grid, 256, L::kTotal, stream>>>(tmA, tmB, p);
  SBM_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

template <int BN, int STAGES, bool kStaged>
static int launch_conv_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiMaps& em,
                            const ConvKernelParams& p, int m_tiles, int n_tiles, int nphase, cudaStream_t stream) {
  using L = Smem2Layout<BN, STAGES, kStaged>;
  static_assert(L::kTotal <= 232448, "shared-memory budget of one CTA exceeded");
  static bool configured = false;
  if (!configured) {
    SBM_CUDA_OK(cudaFuncSetAttribute(conv_igemm_pair_kernel<BN, STAGES, kStaged>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  PairSchedule sch;
  sch.m_tiles = m_tiles;
  sch.m_pairs = (m_tiles + 1) / 2;
  sch.n_tiles = n_tiles;
  sch.nphase = nphase;
  sch.total = nphase * sch.m_pairs * n_tiles;
  const int pairs = std::min(sch.total, sm_count() / 2);
  conv_igemm_pair_kernel<BN, STAGES, kStaged><<<dim3(2 * pairs), 384, L::kTotal, stream>>>(tmA, tmB, em, p, sch);
  SBM_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// Tensor map over an output-geometry tensor (output / residual / bf16 copy) whose box is the 32 rows x kEC columns one
// epilogue warp produces per chunk: dims (c, j, q, i, b); q addresses the output-parity phase of a transposed conv.
static bool encode_rowbox_map(PFN_cuTensorMapEncodeTiled_v12000 encode, CUtensorMap* tm, const void* base, int dtype,
                              int64_t ld, int cout, int ow, int oh, int batch, int sp, int log_ow, int log_th,
                              bool pm = false) {
  const cuuint64_t esz = (dtype == SBM_F32) ? 4 : 2;
  const cuuint64_t OWf = (cuuint64_t)sp * ow, OHf = (cuuint64_t)sp * oh;
  const cuuint64_t dims[5] = {(cuuint64_t)cout, (cuuint64_t)ow, sp == 2 ? OWf + 2 : 1, (cuuint64_t)oh,
                              (cuuint64_t)batch};
  const cuuint64_t strides[4] = {(cuuint64_t)sp * ld * esz, (cuuint64_t)ld * esz, (cuuint64_t)sp * OWf * ld * esz,
                                 OHf * OWf * ld * esz};
  const int th = 1 << log_th;
  const int bw = std::min(ow, 32);
  const int bh = std::min(th, 32 / bw);
  const int bb = 32 / (bw * bh);
  // pixel-major tiles: the warp's 32 rows are 32 samples at one pixel
  const cuuint32_t box[5] = {(cuuint32_t)kEC, pm ? 1u : (cuuint32_t)bw, 1u, pm ? 1u : (cuuint32_t)bh,
                             pm ? 32u : (cuuint32_t)bb};
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  const CUresult cr = encode(tm, dtype == SBM_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                             5, const_cast<void*>(base), dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             dtype == SBM_F32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return cr == CUDA_SUCCESS;
}

static int conv_igemm_impl(const sbm_conv_args* a, cudaStream_t stream) {
  SBM_CHECK_ARG(a != nullptr, "sbm_conv_igemm: null args");
  SBM_CHECK_ARG(a->x && a->wpk && a->out, "sbm_conv_igemm: null operand pointer");
  SBM_CHECK_ARG(a->batch > 0 && a->cin > 0 && a->cout > 0, "sbm_conv_igemm: bad sizes");
  const int lh = ilog2_exact(a->h), lw = ilog2_exact(a->w);
  SBM_CHECK_ARG(lh >= 0 && lw >= 0 && a->h <= 128 && a->w <= 128,
                "sbm_conv_igemm: spatial extent %dx%d must be powers of two <= 128", a->h, a->w);
  SBM_CHECK_ARG(a->ldx % 8 == 0 && a->ldx >= a->cin, "sbm_conv_igemm: ldx=%lld must be a multiple of 8 and >= cin",
                (long long)a->ldx);
  SBM_CHECK_ARG(a->cin_pad % 8 == 0 && a->cin_pad >= a->cin, "sbm_conv_igemm: cin_pad must be a multiple of 8");
  SBM_CHECK_ARG((reinterpret_cast<uintptr_t>(a->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->wpk) & 15) == 0,
                "sbm_conv_igemm: x / wpk must be 16-byte aligned");
  auto encode = get_encode_fn();
  SBM_CHECK_ARG(encode != nullptr, "sbm_conv_igemm: cuTensorMapEncodeTiled entry point not available");

  ConvKernelParams p;
  memset(&p, 0, sizeof(p));
  int oh, ow, nphase = 1;
  // activation view (c, wv, q, hv, b)
  cuuint64_t adim[5];
  cuuint64_t astr[4];
  const int64_t ld = a->ldx;
  if (a->kind == SBM_CONV_S1) {
    SBM_CHECK_ARG(a->kh >= 1 && a->kw >= 1 && (a->kh & 1) && (a->kw & 1) && a->kh * a->kw <= kMaxTaps,
                  "sbm_conv_igemm: stride-1 kernel %dx%d unsupported (odd, <= 16 taps)", a->kh, a->kw);
    oh = a->h; ow = a->w;
    adim[0] = a->cin; adim[1] = a->w; adim[2] = 1; adim[3] = a->h; adim[4] = a->batch;
    astr[0] = ld * 2; astr[1] = (cuuint64_t)a->w * ld * 2; astr[2] = (cuuint64_t)a->w * ld * 2;
    astr[3] = (cuuint64_t)a->h * a->w * ld * 2;
    TapTable& t = p.taps[0];
    const int ph = a->kh / 2, pw = a->kw / 2;
    for (int kh = 0; kh < a->kh; ++kh)
      for (int kw = 0; kw < a->kw; ++kw) {
        const int dh = kh - ph, dw = kw - pw;
        if (abs(dh) >= a->h || abs(dw) >= a->w) continue;  // tap only ever reads zero padding
        t.dh[t.ntaps] = (int8_t)dh; t.dw[t.ntaps] = (int8_t)dw; t.q[t.ntaps] = 0;
        t.wtap[t.ntaps] = (int16_t)(kh * a->kw + kw);
        ++t.ntaps;
      }
  } else if (a->kind == SBM_CONV_S2) {
    SBM_CHECK_ARG((a->kh == 4 && a->kw == 4) || (a->kh == 3 && a->kw == 3),
                  "sbm_conv_igemm: stride-2 kernel must be 4x4 or 3x3");
    SBM_CHECK_ARG(a->h >= 2 && a->w >= 2, "sbm_conv_igemm: stride-2 needs h,w >= 2");
    oh = a->h / 2; ow = a->w / 2;
    adim[0] = a->cin; adim[1] = ow; adim[2] = a->w + 2; adim[3] = oh; adim[4] = a->batch;
    astr[0] = 2 * ld * 2; astr[1] = ld * 2; astr[2] = (cuuint64_t)2 * a->w * ld * 2;
    astr[3] = (cuuint64_t)a->h * a->w * ld * 2;
    TapTable& t = p.taps[0];
    for (int kh = 0; kh < a->kh; ++kh)
      for (int kw = 0; kw < a->kw; ++kw) {
        // input row = 2*oh - 1 + kh = 2*(oh + dh) + parity
        const int rh = kh - 1, rw = kw - 1;
        const int dh = (rh < 0) ? -1 : rh / 2, par_h = (rh < 0) ? 1 : (rh & 1);
        const int dw = (rw < 0) ? -1 : rw / 2, par_w = (rw < 0) ? 1 : (rw & 1);
        if (abs(dh) >= oh && dh != 0) continue;
        if (abs(dw) >= ow && dw != 0) continue;
        t.dh[t.ntaps] = (int8_t)dh; t.dw[t.ntaps] = (int8_t)dw;
        t.q[t.ntaps] = (int16_t)(par_h * a->w + par_w);
        t.wtap[t.ntaps] = (int16_t)(kh * a->kw + kw);
        ++t.ntaps;
      }
  } else if (a->kind == SBM_CONVT_4X4_S2) {
    // ConvTranspose2d(k, stride 2, padding 1) with output 2h x 2w: k = 4 (unet_model.py:30) or k = 3 with
    // output_padding 1 (= the data gradient of the 3x3 stride-2 convolution of unet_openai.py:207)
    SBM_CHECK_ARG((a->kh == 4 && a->kw == 4) || (a->kh == 3 && a->kw == 3),
                  "sbm_conv_igemm: transposed conv must be 4x4 or 3x3");
    SBM_CHECK_ARG(a->h <= 64 && a->w <= 64, "sbm_conv_igemm: transposed conv input must be <= 64x64");
    oh = a->h; ow = a->w;  // per-phase output grid; full output is 2h x 2w
    nphase = 4;
    adim[0] = a->cin; adim[1] = a->w; adim[2] = 1; adim[3] = a->h; adim[4] = a->batch;
    astr[0] = ld * 2; astr[1] = (cuuint64_t)a->w * ld * 2; astr[2] = (cuuint64_t)a->w * ld * 2;
    astr[3] = (cuuint64_t)a->h * a->w * ld * 2;
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw) {
        TapTable& t = p.taps[ph * 2 + pw];
        // out[2i+ph, 2j+pw] += in[i+dh, j+dw] * W[kh][kw] with 2(i+dh) = 2i + ph + 1 - kh
        for (int kh = 0; kh < a->kh; ++kh) {
          if (((ph + 1 - kh) & 1) != 0) continue;
          const int dh = (ph + 1 - kh) / 2;
          for (int kw = 0; kw < a->kw; ++kw) {
            if (((pw + 1 - kw) & 1) != 0) continue;
            const int dw = (pw + 1 - kw) / 2;
            if (abs(dh) >= a->h || abs(dw) >= a->w) continue;
            t.dh[t.ntaps] = (int8_t)dh; t.dw[t.ntaps] = (int8_t)dw; t.q[t.ntaps] = 0;
            t.wtap[t.ntaps] = (int16_t)(kh * a->kw + kw);
            ++t.ntaps;
          }
        }
      }
  } else {
    SBM_CHECK_ARG(false, "sbm_conv_igemm: unknown kind %d", a->kind);
  }

  const int log_oh = ilog2_exact(oh), log_ow = ilog2_exact(ow);
  const int log_ohw = log_oh + log_ow;
  const int log_th = (log_ohw >= 7) ? (7 - log_ow) : log_oh;
  const int nb = (log_ohw >= 7) ? 1 : (1 << (7 - log_ohw));
  p.batch = a->batch;
  p.log_ow = log_ow; p.log_th = log_th; p.log_oh = log_oh;
  p.cin = a->cin; p.cout = a->cout;
  p.cblocks = (a->cin + kBK - 1) / kBK;
  p.bias = a->bias; p.residual = a->residual; p.out = a->out; p.out2 = a->out2; p.stats = a->stats;
  p.act = a->act; p.out_dtype = a->out_dtype; p.res_dtype = a->res_dtype;
  p.out2_preact = (a->out2 != nullptr && a->out2_preact) ? 1 : 0;
  p.rowbias = a->rowbias; p.ld_rowbias = a->ld_rowbias;
  p.bias_vec = (a->bias != nullptr && (reinterpret_cast<uintptr_t>(a->bias) & 15) == 0) ? 1 : 0;
  if (a->gn_tab != nullptr) {
    SBM_CHECK_ARG(a->kind == SBM_CONV_S1 && (a->kh == 1 || a->kh == 3) && a->kh == a->kw,
                  "sbm_conv_igemm: GroupNorm folding supports 1x1 and 3x3 stride-1 convolutions");
    SBM_CHECK_ARG(a->gn_stats != nullptr && a->gn_count > 0 && a->bias == nullptr,
                  "sbm_conv_igemm: folded GroupNorm needs input statistics and carries the bias in its table");
    p.gn_stats = a->gn_stats; p.gn_tab = a->gn_tab;
    p.gn_inv_count = 1.0 / (double)a->gn_count; p.gn_eps = a->gn_eps;
  }

  // output addressing
  const int64_t OHf = (a->kind == SBM_CONVT_4X4_S2) ? 2 * oh : oh;
  const int64_t OWf = (a->kind == SBM_CONVT_4X4_S2) ? 2 * ow : ow;
  const int64_t sp = (a->kind == SBM_CONVT_4X4_S2) ? 2 : 1;  // spatial step between a phase's neighbours
  if (a->out_nchw) {
    SBM_CHECK_ARG(a->out_dtype == SBM_F32 && a->residual == nullptr && a->out2 == nullptr,
                  "sbm_conv_igemm: NCHW output is fp32 without residual");
    p.o_sc = OHf * OWf; p.o_sw = sp; p.o_sh = sp * OWf; p.o_sb = (int64_t)a->cout * OHf * OWf;
  } else {
    SBM_CHECK_ARG(a->ldo >= a->cout, "sbm_conv_igemm: ldo < cout");
    p.o_sc = 1; p.o_sw = sp * a->ldo; p.o_sh = sp * OWf * a->ldo; p.o_sb = OHf * OWf * a->ldo;
  }
  p.r_sw = sp * a->ldr; p.r_sh = sp * OWf * a->ldr; p.r_sb = OHf * OWf * a->ldr;
  p.o2_sw = sp * a->ldo2; p.o2_sh = sp * OWf * a->ldo2; p.o2_sb = OHf * OWf * a->ldo2;
  for (int ph = 0; ph < nphase; ++ph) {
    const int64_t phh = ph >> 1, pww = ph & 1;
    int64_t off = 0, off2 = 0;
    if (a->kind == SBM_CONVT_4X4_S2) {
      off = a->out_nchw ? (phh * OWf + pww) : (phh * OWf + pww) * a->ldo;
      off2 = (phh * OWf + pww) * a->ldo2;
      SBM_CHECK_ARG(a->residual == nullptr, "sbm_conv_igemm: transposed conv does not take a residual");
    }
    SBM_CHECK_ARG(off < (int64_t(1) << 31) && off2 < (int64_t(1) << 31), "sbm_conv_igemm: phase offset overflow");
    p.taps[ph].out_off = (int32_t)off;
    p.taps[ph].out2_off = (int32_t)off2;
    SBM_CHECK_ARG(p.taps[ph].ntaps > 0, "sbm_conv_igemm: empty tap table");
  }
  // 16-byte vector access is legal when every row start and channel chunk is 16-byte aligned
  const int esz = (a->out_dtype == SBM_F32) ? 4 : 2;
  bool vec = !a->out_nchw && ((a->ldo * esz) % 16 == 0) && ((reinterpret_cast<uintptr_t>(a->out) & 15) == 0);
  if (a->residual) {
    const int rsz = (a->res_dtype == SBM_F32) ? 4 : 2;
    vec = vec && ((a->ldr * rsz) % 16 == 0) && ((reinterpret_cast<uintptr_t>(a->residual) & 15) == 0);
  }
  if (a->out2) vec = vec && ((a->ldo2 * 2) % 16 == 0) && ((reinterpret_cast<uintptr_t>(a->out2) & 15) == 0);
  p.vec_ok = vec ? 1 : 0;

  // ---- tile shape
  const int64_t M = (int64_t)a->batch << log_ohw;
  static const int kNarrowPct = [] { const char* e = getenv("SBM_NARROW_PCT"); return e ? atoi(e) : 75; }();
  static const int kPairPct = [] { const char* e = getenv("SBM_PAIR_PCT"); return e ? atoi(e) : 80; }();
  int BN = 0;
  auto choose = [&](int m_tiles_) -> bool {  // sets BN; returns whether the CTA-pair kernel runs this tile list
    if (a->cout <= 32) BN = 32;
    else if (a->cout <= 64) BN = 64;
    else if (a->cout <= 128) BN = 128;
    else if (a->cout % 256 == 0 || a->cout > 512) BN = 256;
    else BN = (a->cout % 128 == 0) ? 128 : 256;
    // small problems (one tile per CTA, less than a wave): narrower tiles spread the K loop and the epilogue over more
    // SMs (stop at kNarrowPct % of a wave: a 128-wide tile list that covers 86 % of the SMs beats 64-wide tiles in 1.7
    // waves, whose A tiles are re-read from L2 by twice as many CTAs)
    while (BN > 64 && (int64_t)m_tiles_ * ((a->cout + BN - 1) / BN) * nphase * 100 < (int64_t)sm_count() * kNarrowPct)
      BN >>= 1;
    // CTA-pair kernel: 256 x BN tiles; worth it once the tile pairs cover ~40 % of the SM pairs (measured on the
    // 2x2-level layers: 64 pair tiles on 74 SM pairs run 1.3x faster than 128 single-CTA tiles -- half the weight
    // traffic per CTA).  A cout tail is fine: weight rows beyond cout are TMA zero fill, the epilogue clips the columns
    // BN == 64 (cout <= 64: the PolyMNIST net's 42- and 64-channel layers): the persistent pair kernel only pays once
    // the list is long -- a single-CTA launch of one tile per CTA has no epilogue / main-loop overlap, and at 64k
    // latents these layers ran at a tenth of the tensor peak (round-1 sweep)
    static const int kPair64 = [] { const char* e = getenv("SBM_PAIR_BN64"); return e ? atoi(e) : 1; }();
    const bool bn_ok = BN == 256 || BN == 128 || (BN == 64 && kPair64 && a->cout > 32);
    return !g_force_single && bn_ok && !a->out_nchw &&
           (int64_t)((m_tiles_ + 1) / 2) * ((a->cout + BN - 1) / BN) * nphase * 200 >= (int64_t)sm_count() * kPairPct;
  };
  // Pixel-major tiling (see ConvKernelParams::pm): a 'same' convolution computes taps on zero padding for every border
  // pixel -- 8 % of the MMAs of a 3x3 at 16x16, 16 % at 8x8, 31 % at 4x4, 56 % at 2x2.  With the 128 rows of a tile
  // taken along the BATCH at one output pixel, those taps are skipped for the whole tile (and never loaded).  Used when
  // the saving outweighs the rows wasted by rounding the batch up to 256-sample blocks.
  int m_tiles = (int)((M + kBM - 1) / kBM);
  bool pm = false;
  if (a->kind == SBM_CONV_S1 && p.taps[0].ntaps > 1 && g_pixel_major != 0) {
    int64_t valid = 0;
    for (int i = 0; i < oh; ++i)
      for (int j = 0; j < ow; ++j)
        for (int k = 0; k < p.taps[0].ntaps; ++k)
          valid += (i + p.taps[0].dh[k] >= 0 && i + p.taps[0].dh[k] < oh && j + p.taps[0].dw[k] >= 0 &&
                    j + p.taps[0].dw[k] < ow) ? 1 : 0;
    const int blocks = (a->batch + 255) / 256;
    const double work_pm = (double)blocks * 256 * valid;                            // MMA rows x taps
    const double work_std = (double)m_tiles * kBM * p.taps[0].ntaps;
    if (g_pixel_major == 1 || work_pm < 0.97 * work_std) {
      const int m_tiles_pm = 2 * blocks * oh * ow;
      if (oh * ow <= 256 && choose(m_tiles_pm)) {
        pm = true;
        m_tiles = m_tiles_pm;
        p.pm_blocks = blocks;
        // few rounds per cluster: balance matters most -> global cost order; many rounds: keep a block's pixels together
        p.pm_global = ((int64_t)(m_tiles_pm / 2) * ((a->cout + BN - 1) / BN) <= 4 * (int64_t)(sm_count() / 2)) ? 1 : 0;
        if (const char* e = getenv("SBM_PM_GLOBAL")) p.pm_global = atoi(e);
        int n = 0;
        for (int want = p.taps[0].ntaps; want >= 1; --want)     // counting sort by valid taps, raster order inside
          for (int i = 0; i < oh; ++i)
            for (int j = 0; j < ow; ++j) {
              int v = 0;
              for (int k = 0; k < p.taps[0].ntaps; ++k)
                v += (i + p.taps[0].dh[k] >= 0 && i + p.taps[0].dh[k] < oh && j + p.taps[0].dw[k] >= 0 &&
                      j + p.taps[0].dw[k] < ow) ? 1 : 0;
              if (v == want) p.pm_pix[n++] = (uint8_t)(i * ow + j);
            }
      }
    }
  }
  const bool use_pair = pm ? true : choose(m_tiles);
  const int n_tiles_pair = (a->cout + BN - 1) / BN;
  p.pm = pm ? 1 : 0;

  // ---- tensor maps
  CUtensorMap tmA, tmB;
  const cuuint32_t abox_std[5] = {(cuuint32_t)kBK, (cuuint32_t)ow, 1u, (cuuint32_t)(1 << log_th), (cuuint32_t)nb};
  const cuuint32_t abox_pm[5] = {(cuuint32_t)kBK, 1u, 1u, 1u, (cuuint32_t)kBM};
  const cuuint32_t* abox = pm ? abox_pm : abox_std;
  const cuuint32_t ones5[5] = {1, 1, 1, 1, 1};
  CUresult cr = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(a->x), adim, astr, abox, ones5,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SBM_CHECK_ARG(cr == CUDA_SUCCESS, "sbm_conv_igemm: activation tensor map encode failed (CUresult %d)", (int)cr);

  const int ntaps_total = a->kh * a->kw;
  const cuuint64_t bdim[3] = {(cuuint64_t)a->cin, (cuuint64_t)a->cout, (cuuint64_t)ntaps_total};
  const cuuint64_t bstr[2] = {(cuuint64_t)a->cin_pad * 2, (cuuint64_t)a->cout * a->cin_pad * 2};
  const cuuint32_t bbox[3] = {(cuuint32_t)kBK, (cuuint32_t)(use_pair ? BN / 2 : BN), 1u};
  const cuuint32_t ones3[3] = {1, 1, 1};
  cr = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(a->wpk), bdim, bstr, bbox, ones3,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SBM_CHECK_ARG(cr == CUDA_SUCCESS, "sbm_conv_igemm: weight tensor map encode failed (CUresult %d)", (int)cr);

  if (use_pair) {
    // staged epilogue: TMA box stores / residual loads; needs 16-byte aligned rows (vec) and at least one full warp box
    EpiMaps em;
    bool staged = !g_force_direct && vec && M >= 32;
    if (staged) {
      const int sp_i = (int)sp;
      staged = encode_rowbox_map(encode, &em.out, a->out, a->out_dtype, a->ldo, a->cout, ow, oh, a->batch, sp_i, log_ow,
                                 log_th, pm);
      if (staged && a->residual)
        staged = encode_rowbox_map(encode, &em.res, a->residual, a->res_dtype, a->ldr, a->cout, ow, oh, a->batch, sp_i,
                                   log_ow, log_th, pm);
      if (staged && a->out2)
        staged = encode_rowbox_map(encode, &em.out2, a->out2, SBM_BF16, a->ldo2, a->cout, ow, oh, a->batch, sp_i,
                                   log_ow, log_th, pm);
    }
    g_last_variant = BN | (1 << 16) | (staged ? (1 << 17) : 0) | (pm ? (1 << 18) : 0);
    if (staged) {
      if (!a->residual) em.res = em.out;
      if (!a->out2) em.out2 = em.out;
      for (int ph = 0; ph < nphase; ++ph)
        p.taps[ph].out_q = (a->kind == SBM_CONVT_4X4_S2) ? (int32_t)((ph >> 1) * OWf + (ph & 1)) : 0;
      if (BN == 256) return launch_conv_pair<256, 5, true>(tmA, tmB, em, p, m_tiles, n_tiles_pair, nphase, stream);
      if (BN == 128) return launch_conv_pair<128, 6, true>(tmA, tmB, em, p, m_tiles, n_tiles_pair, nphase, stream);
      return launch_conv_pair<64, 8, true>(tmA, tmB, em, p, m_tiles, n_tiles_pair, nphase, stream);
    }
    memset(&em, 0, sizeof(em));
    if (BN == 256) return launch_conv_pair<256, 6, false>(tmA, tmB, em, p, m_tiles, n_tiles_pair, nphase, stream);
    if (BN == 128) return launch_conv_pair<128, 8, false>(tmA, tmB, em, p, m_tiles, n_tiles_pair, nphase, stream);
    return launch_conv_pair<64, 10, false>(tmA, tmB, em, p, m_tiles, n_tiles_pair, nphase, stream);
  }
  dim3 grid((unsigned)m_tiles, (unsigned)((a->cout + BN - 1) / BN), (unsigned)nphase);
  g_last_variant = BN;
  switch (BN) {
    case 32: return launch_conv<32, 10>(tmA, tmB, p, grid, stream);  // small tiles: deeper ring, the K loop is TMA-latency bound
    case 64: return launch_conv<64, 8>(tmA, tmB, p, grid, stream);
    case 128: return launch_conv<128, 6>(tmA, tmB, p, grid, stream);
    default: return launch_conv<256, 4>(tmA, tmB, p, grid, stream);
  }
}

// ---------------------------------------------------------------- weight packing
__global__ void pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int taps, int rows,
                                   int cols, int cols_pad, int64_t s_tap, int64_t s_row, int64_t s_col) {
  const int64_t total = (int64_t)taps * rows * cols_pad;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cols_pad);
    const int64_t tr = idx / cols_pad;
    const int r = (int)(tr % rows);
    const int t = (int)(tr / rows);
    float v = 0.f;
    if (c < cols) v = w[t * s_tap + r * s_row + c * s_col];
    dst[idx] = __float2bfloat16_rn(v);
  }
}

// Tiled variant: a block re-lays one 32 x 32 (rows x cols) tile for all taps through shared memory, reading in
// SOURCE-contiguous order (taps, then whichever of row / col has the smaller stride) and writing runs of 32 bf16 along
// the destination's column axis -- both sides coalesced (the element-wise kernel above reads conv weights with a
// stride of kh*kw floats; re-packing 223 M parameters every training step made that 7 % of the step).
struct PackDesc {          // one weight tensor of a multi-tensor re-pack
  const float* src;
  __nv_bfloat16* dst;
  int32_t taps, rows, cols, cols_pad;
  int64_t s_tap, s_row, s_col;
  int32_t tiles_c, first_block;   // 32-column tiles per tile row; index of this tensor's first block
  uint32_t taps_magic;
  int32_t pad;
};
static_assert(sizeof(PackDesc) == sizeof(sbm_pack_desc), "PackDesc must mirror sbm_pack_desc");

__device__ __forceinline__ void pack_tile(const PackDesc& d, int tile, float* tile_smem) {
  const int taps = d.taps, rows = d.rows, cols = d.cols, cols_pad = d.cols_pad;
  const int TP = taps | 1;  // [32 rows][33][TP]: conflict-free for both access orders
  const int r0 = (tile / d.tiles_c) * 32, c0 = (tile % d.tiles_c) * 32;
  const bool col_inner = llabs(d.s_col) <= llabs(d.s_row);
  const int n = taps * 1024;
  // Fast path: weights whose taps are contiguous in memory (|s_tap| == 1) and whose faster tensor index has stride
  // `taps` -- the nn.Conv2d layout [O][I][taps] read as (rows = O, cols = I) for the forward operand, or as
  // (rows = I, cols = O) with the taps walked backwards for the data-gradient operand.  For a fixed index of the slower
  // ("major") dimension the 32 faster ("minor") indices x taps form ONE contiguous run of 32 * taps floats: 16-byte
  // loads, issued in batches of up to 9 per thread before any is consumed (the kernel is memory-latency bound: ncu,
  // round 2).  The generic path below reads the data-gradient operands (half of all packs of a training step) with
  // one 4-byte load in flight per thread.
  const bool tap_fwd = d.s_tap == 1, tap_rev = d.s_tap == -1;
  const bool col_minor = d.s_col == taps, row_minor = !col_minor && d.s_row == taps;
  const int64_t s_major = col_minor ? d.s_row : d.s_col;
  const float* run0 = d.src - (tap_rev ? taps - 1 : 0);
  if ((tap_fwd || tap_rev) && (col_minor || row_minor) && (taps & 1) && taps > 1 && (s_major & 3) == 0 &&
      (col_minor ? c0 + 32 <= cols : r0 + 32 <= rows) && (reinterpret_cast<uintptr_t>(run0) & 15) == 0) {
    const int run4 = taps * 8;  // float4s per run
    const int minor0 = col_minor ? c0 : r0, major0 = col_minor ? r0 : c0, major_n = col_minor ? rows : cols;
    for (int base = 0; base < 32 * run4; base += 9 * 256) {
      float4 v[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int e = base + i * 256 + (int)threadIdx.x;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < 32 * run4) {
          const int mj = e / run4, j4 = e - mj * run4;
          if (major0 + mj < major_n)
            v[i] = __ldg(reinterpret_cast<const float4*>(run0 + (int64_t)(major0 + mj) * s_major + (int64_t)minor0 * taps) + j4);
        }
      }
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int e = base + i * 256 + (int)threadIdx.x;
        if (e < 32 * run4) {
          const int mj = e / run4, j4 = e - mj * run4;
          const float vv[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int k = 4 * j4 + u;
            const int mi = (int)__umulhi((uint32_t)k, d.taps_magic);   // k / taps
            const int tt = k - mi * taps;
            const int t = tap_rev ? taps - 1 - tt : tt;
            const int rr = col_minor ? mj : mi, cc = col_minor ? mi : mj;
            tile_smem[(rr * 33 + cc) * TP + t] = vv[u];
          }
        }
      }
    }
  } else {
#pragma unroll 4
    for (int e = threadIdx.x; e < n; e += 256) {
      const int rc = taps == 1 ? e : (int)__umulhi((uint32_t)e, d.taps_magic);  // e / taps (exact for e < 2^16)
      const int t = e - rc * taps;
      const int inner = rc & 31, outer = rc >> 5;
      const int rr = col_inner ? outer : inner, cc = col_inner ? inner : outer;
      const int r = r0 + rr, c = c0 + cc;
      float v = 0.f;
      if (r < rows && c < cols) v = __ldg(d.src + t * d.s_tap + (int64_t)r * d.s_row + (int64_t)c * d.s_col);
      tile_smem[(rr * 33 + cc) * TP + t] = v;
    }
  }
  __syncthreads();
  if ((cols_pad & 7) == 0 && (reinterpret_cast<uintptr_t>(d.dst) & 15) == 0) {
    // 8 consecutive columns (one 16-byte store) per thread
    for (int e = threadIdx.x; e < taps * 128; e += 256) {
      const int c8 = e & 3, rr = (e >> 2) & 31, t = e >> 7;
      const int r = r0 + rr, c = c0 + c8 * 8;
      if (r < rows && c < cols_pad) {
        uint32_t wds[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float lo = tile_smem[(rr * 33 + c8 * 8 + 2 * k) * TP + t], hi = tile_smem[(rr * 33 + c8 * 8 + 2 * k + 1) * TP + t];
          __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
          wds[k] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(d.dst + ((int64_t)t * rows + r) * cols_pad + c) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
      }
    }
  } else {
#pragma unroll 4
    for (int e = threadIdx.x; e < n; e += 256) {
      const int cc = e & 31, rr = (e >> 5) & 31, t = e >> 10;
      const int r = r0 + rr, c = c0 + cc;
      if (r < rows && c < cols_pad)
        d.dst[((int64_t)t * rows + r) * cols_pad + c] = __float2bfloat16_rn(tile_smem[(rr * 33 + cc) * TP + t]);
    }
  }
}

__global__ void __launch_bounds__(256)
pack_weight_tiled_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int taps, int rows, int cols,
                         int cols_pad, int64_t s_tap, int64_t s_row, int64_t s_col, uint32_t taps_magic) {
  extern __shared__ float tile[];
  PackDesc d;
  d.src = w; d.dst = dst; d.taps = taps; d.rows = rows; d.cols = cols; d.cols_pad = cols_pad;
  d.s_tap = s_tap; d.s_row = s_row; d.s_col = s_col; d.tiles_c = gridDim.x; d.first_block = 0; d.taps_magic = taps_magic;
  pack_tile(d, blockIdx.y * gridDim.x + blockIdx.x, tile);
}

// all stale weight packs of a net in ONE launch: block -> (descriptor, tile) by binary search over first_block
__global__ void __launch_bounds__(256)
pack_weights_multi_kernel(const PackDesc* __restrict__ descs, int n_descs) {
  extern __shared__ float tile[];
  int lo = 0, hi = n_descs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].first_block <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const PackDesc d = descs[lo];
  pack_tile(d, (int)blockIdx.x - d.first_block, tile);
}

// One block per output channel n: writes the gamma-folded bf16 weight rows [tap][n][:] and the per-tap partial sums
//   Pg[tap] = sum_c bf16(w*gamma)   Pb[tap] = sum_c w*beta
// then combines them into the 16 border classes (which of the 3x3 taps see real pixels).
__global__ void __launch_bounds__(256)
fold_groupnorm_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int kh, int kw, int rows, int cols,
                      int cols_pad, int64_t s_tap, int64_t s_row, int64_t s_col, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const float* __restrict__ bias, float* __restrict__ tab) {
  __shared__ double red[2][9][8];
  __shared__ double tot[2][9];
  const int n = blockIdx.x;
  const int taps = kh * kw;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int t = 0; t < taps; ++t) {
    float pg = 0.f, pb = 0.f;
    for (int c = threadIdx.x; c < cols_pad; c += blockDim.x) {
      float wg = 0.f;
      if (c < cols) {
        const float wv = w[t * s_tap + n * s_row + c * s_col];
        const __nv_bfloat16 h = __float2bfloat16_rn(wv * __ldg(gamma + c));
        wg = __bfloat162float(h);
        pb += wv * __ldg(beta + c);
        dst[((int64_t)t * rows + n) * cols_pad + c] = h;
      } else {
        dst[((int64_t)t * rows + n) * cols_pad + c] = __float2bfloat16_rn(0.f);
      }
      pg += wg;
    }
    const double dg = warp_sum((double)pg), db = warp_sum((double)pb);
    if (lane == 0) { red[0][t][wid] = dg; red[1][t][wid] = db; }
  }
  __syncthreads();
  if (threadIdx.x < 2 * taps) {
    const int which = threadIdx.x / taps, t = threadIdx.x % taps;
    double v = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) v += red[which][t][i];
    tot[which][t] = v;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int which = threadIdx.x >> 4, cls = threadIdx.x & 15;
    double v = 0.0;
    for (int a = 0; a < kh; ++a)
      for (int b = 0; b < kw; ++b) {
        const bool okh = (kh == 1) || a == 1 || (a == 0 && (cls & 1)) || (a == 2 && (cls & 2));
        const bool okw = (kw == 1) || b == 1 || (b == 0 && (cls & 4)) || (b == 2 && (cls & 8));
        if (okh && okw) v += tot[which][a * kw + b];
      }
    if (which == 1 && bias != nullptr) v += (double)bias[n];
    tab[((int64_t)which * 16 + cls) * rows + n] = (float)v;
  }
}

}  // namespace sbm

// ------------------------------------------------------------------------------------ C ABI
static thread_local char g_err[512] = "";
void sbm::set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {

const char* sbm_last_error(void) { return g_err; }
int sbm_version(void) { return 100; }
unsigned long long sbm_launch_count(void) { return sbm::g_launches.load(); }

int sbm_conv_force_single_cta(int32_t on) {
  sbm::g_force_single = on != 0;
  return 0;
}

int sbm_conv_last_variant(void) { return sbm::g_last_variant; }
int sbm_conv_pixel_major(int32_t mode) {
  const int old = sbm::g_pixel_major;
  sbm::g_pixel_major = mode < 0 ? -1 : (mode > 0 ? 1 : 0);
  return old;
}

int sbm_conv_force_direct_epilogue(int32_t on) {
  sbm::g_force_direct = on != 0;
  return 0;
}

int sbm_conv_igemm(const sbm_conv_args* a, void* stream) {
  return sbm::conv_igemm_impl(a, static_cast<cudaStream_t>(stream));
}

int sbm_conv_fold_groupnorm(const float* w, void* dst, float* tab, int32_t kh, int32_t kw, int32_t rows, int32_t cols,
                            int32_t cols_pad, int64_t s_tap, int64_t s_row, int64_t s_col, const float* gamma,
                            const float* beta, const float* bias, void* stream) {
  SBM_CHECK_ARG(w && dst && tab && gamma && beta && rows > 0 && cols > 0 && cols_pad >= cols,
                "sbm_conv_fold_groupnorm: bad args");
  SBM_CHECK_ARG(kh == kw && (kh == 1 || kh == 3), "sbm_conv_fold_groupnorm: 1x1 or 3x3 kernels only");
  sbm::fold_groupnorm_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, static_cast<__nv_bfloat16*>(dst), kh, kw, rows, cols, cols_pad, s_tap, s_row, s_col, gamma, beta, bias, tab);
  SBM_CUDA_OK(cudaGetLastError());
  sbm::g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int sbm_pack_weights_multi(const sbm_pack_desc* descs_dev, int32_t n_descs, int32_t n_blocks, int32_t max_taps,
                           void* stream) {
  SBM_CHECK_ARG(descs_dev && n_descs > 0 && n_blocks > 0 && max_taps > 0 && max_taps <= 64,
                "sbm_pack_weights_multi: bad args");
  const size_t smem = (size_t)(max_taps | 1) * 32 * 33 * sizeof(float);
  SBM_CHECK_ARG(smem <= 200 * 1024, "sbm_pack_weights_multi: too many taps");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    SBM_CUDA_OK(cudaFuncSetAttribute(sbm::pack_weights_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    configured = smem;
  }
  sbm::pack_weights_multi_kernel<<<n_blocks, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const sbm::PackDesc*>(descs_dev), n_descs);
  SBM_CUDA_OK(cudaGetLastError());
  sbm::g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int sbm_pack_weight_bf16(const float* w, void* dst, int32_t taps, int32_t rows, int32_t cols, int32_t cols_pad,
                         int64_t s_tap, int64_t s_row, int64_t s_col, void* stream) {
  SBM_CHECK_ARG(w && dst && taps > 0 && rows > 0 && cols > 0 && cols_pad >= cols, "sbm_pack_weight_bf16: bad args");
  if ((int64_t)rows * cols >= 4096 && taps <= 16) {
    const size_t smem = (size_t)(taps | 1) * 32 * 33 * sizeof(float);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      SBM_CUDA_OK(cudaFuncSetAttribute(sbm::pack_weight_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
      configured = smem;
    }
    dim3 grid((cols_pad + 31) / 32, (rows + 31) / 32);
    sbm::pack_weight_tiled_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
        w, static_cast<__nv_bfloat16*>(dst), taps, rows, cols, cols_pad, s_tap, s_row, s_col,
        (uint32_t)(taps > 1 ? ((1ull << 32) + taps - 1) / taps : 0));
  } else {
    const int64_t total = (int64_t)taps * rows * cols_pad;
    const int threads = 256;
    const int blocks = (int)std::min<int64_t>((total + threads - 1) / threads, (int64_t)sbm::sm_count() * 8);
    sbm::pack_weight_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        w, static_cast<__nv_bfloat16*>(dst), taps, rows, cols, cols_pad, s_tap, s_row, s_col);
  }
  SBM_CUDA_OK(cudaGetLastError());
  sbm::g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // extern "C"
