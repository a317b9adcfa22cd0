// CTA-pair (cta_group::2) implicit-GEMM convolution kernel, its staged TMA epilogue and the launch template.
// Included by conv_igemm.cu (host side, run-time tested epilogue) and by conv_pair_modes*.cu (one kernel
// instantiation per statically compiled epilogue mode).  See conv_igemm.cu for the algorithm description.
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <algorithm>
#include <atomic>

#include "../../include/sbmae_b200.h"
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace sbm {

extern std::atomic<unsigned long long> g_launches;

// Pipeline trace of cluster 0 (compile with -DSBM_PAIR_TRACE, tools/trace_pair.py): one record per event
#ifdef SBM_PAIR_TRACE
#define SBM_TRACE(role, ev, round)                                                                              \
  do {                                                                                                           \
    if (p.trace != nullptr && blockIdx.x < 2) {                                                                  \
      const unsigned long long i_ = atomicAdd(p.trace, 1ull);                                                    \
      if (i_ < 60000) p.trace[1 + i_] = (ptx::globaltimer_ns() << 20) | ((unsigned long long)(blockIdx.x & 1) << 19) | ((unsigned long long)(role) << 16) | ((unsigned long long)(ev) << 12) | (unsigned long long)((round) & 0xFFF); \
    }                                                                                                            \
  } while (0)
#else
#define SBM_TRACE(role, ev, round) do {} while (0)
#endif

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
  int32_t rowbias_vec;      // rows of `rowbias` are 16-byte aligned
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
#ifdef SBM_PAIR_TRACE
  unsigned long long* trace;   // [0] = counter, then (globaltimer << 20 | role << 16 | event << 12 | round) records of cluster 0
#endif
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
// The staging buffers are addressed in the shared state space explicitly (32-bit addresses, st.shared / ld.shared):
// through generic pointers the compiler emitted ST.E.128 / LD.E.128, whose completion the proxy fence before every
// TMA store then had to wait for -- more than half of the epilogue time of the K-short layers (round-2 measurement).
__device__ __forceinline__ uint32_t stg_f32(uint32_t buf, int r, int k) {
  return buf + r * 64 + ((k ^ ((r >> 1) & 3)) << 4);
}
__device__ __forceinline__ uint32_t stg_bf16(uint32_t buf, int r, int k) {
  return buf + r * 32 + ((k ^ ((r >> 2) & 1)) << 4);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void stg_store_bf16_row(uint32_t buf, int r, const float* f) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat162 t = __floats2bfloat162_rn(f[8 * k + 2 * e], f[8 * k + 2 * e + 1]);
      w[e] = *reinterpret_cast<uint32_t*>(&t);
    }
    sts128(stg_bf16(buf, r, k), w[0], w[1], w[2], w[3]);
  }
}
// ---- epilogue modes.  The staged chunk loop tested ~40 kernel parameters per 16-column chunk (constant load ->
// uniform compare -> branch, each a short-scoreboard + branch-resolve stall): ncu's source view of a K-short layer
// showed the epilogue warps spending their time on exactly those, not on TMEM / shared memory / TMA, and a pure-store
// 1x1 convolution ran at 2.4 TB/s whatever it wrote.  The flag combinations the score nets use are therefore compiled
// as straight-line loops (template MODE = a bit set below, selected ONCE per launch); anything else takes EM_DYN, the
// old run-time tested loop.
enum : uint32_t {
  EM_GN = 1u, EM_BIAS = 2u, EM_ROWB = 4u, EM_GELU = 8u, EM_SILU = 16u, EM_RES32 = 32u, EM_RES16 = 64u, EM_OBF16 = 128u,
  EM_O2 = 256u, EM_O2PRE = 512u, EM_STATS = 1024u, EM_DYN = 1u << 31,
  // split-K work items: no epilogue arithmetic, the raw fp32 accumulators of K slice s go to slab s of a workspace by
  // TMA box stores (a second kernel sums the slabs in slice order and applies the epilogue: deterministic); the tile
  // list carries a K-slice index
  EM_SPLITK = 1u << 17
};
// EM_DYN: the same bits, computed from the kernel parameters ONCE per tile into a register (`fl`), so that the chunk
// loop tests register predicates instead of re-loading and comparing constant-bank parameters.
constexpr uint32_t EM_BIASVEC = 1u << 16;  // run-time mode only: bias is 16-byte aligned
__device__ __forceinline__ uint32_t epi_flags(const ConvKernelParams& p) {
  uint32_t fl = 0;
  if (p.gn_tab != nullptr) fl |= EM_GN;
  if (p.bias != nullptr) fl |= EM_BIAS;
  if (p.bias_vec) fl |= EM_BIASVEC;
  if (p.rowbias != nullptr) fl |= EM_ROWB;
  if (p.act == SBM_ACT_GELU) fl |= EM_GELU;
  if (p.act == SBM_ACT_SILU) fl |= EM_SILU;
  if (p.residual != nullptr) fl |= (p.res_dtype == SBM_F32) ? EM_RES32 : EM_RES16;
  if (p.out_dtype == SBM_BF16) fl |= EM_OBF16;
  if (p.out2 != nullptr) fl |= p.out2_preact ? EM_O2PRE : EM_O2;
  if (p.stats != nullptr) fl |= EM_STATS;
  return fl;
}
template <uint32_t MODE>
struct EpiMode {
  static constexpr bool dyn = (MODE & EM_DYN) != 0;
  // static modes imply 16-byte aligned bias / GroupNorm table rows (cout % 4 == 0 when a table is present)
  __device__ __forceinline__ static bool has(uint32_t fl, uint32_t bits) { return ((dyn ? fl : MODE) & bits) != 0; }
  __device__ __forceinline__ static bool gn(uint32_t fl) { return has(fl, EM_GN); }
  __device__ __forceinline__ static bool bias(uint32_t fl) { return has(fl, EM_BIAS); }
  __device__ __forceinline__ static bool rowb(uint32_t fl) { return has(fl, EM_ROWB); }
  __device__ __forceinline__ static bool gelu(uint32_t fl) { return has(fl, EM_GELU); }
  __device__ __forceinline__ static bool silu_(uint32_t fl) { return has(fl, EM_SILU); }
  __device__ __forceinline__ static bool res(uint32_t fl) { return has(fl, EM_RES32 | EM_RES16); }
  __device__ __forceinline__ static bool res32(uint32_t fl) { return has(fl, EM_RES32); }
  __device__ __forceinline__ static bool obf16(uint32_t fl) { return has(fl, EM_OBF16); }
  __device__ __forceinline__ static bool o2(uint32_t fl) { return has(fl, EM_O2 | EM_O2PRE); }
  __device__ __forceinline__ static bool o2pre(uint32_t fl) { return has(fl, EM_O2PRE); }
  __device__ __forceinline__ static bool stats(uint32_t fl) { return has(fl, EM_STATS); }
  // a column tail (cout % 16 != 0: the 170-channel stem) takes the clamped scalar loads in its last chunk only
  __device__ __forceinline__ static bool full(int n, int cout) { return n + 16 <= cout; }
  __device__ __forceinline__ static bool vec(int n, int cout) { return n + 16 <= cout && (!dyn || (cout & 3) == 0); }
  __device__ __forceinline__ static bool biasvec(uint32_t fl, int n, int cout) {
    return n + 16 <= cout && (!dyn || (fl & EM_BIASVEC) != 0);
  }
};

// `stage` holds the residual chunk on entry (when there is one) and the output chunk on exit; `stage2` receives the
// bf16 copy.  v = the 16 fp32 accumulators of this thread's row.
template <uint32_t MODE>
__device__ __forceinline__ void epilogue_chunk_staged(const ConvKernelParams& p, const uint32_t* v, int n, bool row_ok,
                                                      int b, int lane, uint32_t stage, uint32_t stage2,
                                                      float& s1, float& s2, const GnRow& gr, uint32_t fl) {
  using M = EpiMode<MODE>;
  float f[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]);
  const int cmax = p.cout - 1;
  if (M::gn(fl)) {
    if (M::vec(n, p.cout)) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(gr.sg + n) + k);
        const float4 t = __ldg(reinterpret_cast<const float4*>(gr.tb + n) + k);
        f[4 * k] = fmaf(gr.rstd, f[4 * k] - gr.mu * a.x, t.x);
        f[4 * k + 1] = fmaf(gr.rstd, f[4 * k + 1] - gr.mu * a.y, t.y);
        f[4 * k + 2] = fmaf(gr.rstd, f[4 * k + 2] - gr.mu * a.z, t.z);
        f[4 * k + 3] = fmaf(gr.rstd, f[4 * k + 3] - gr.mu * a.w, t.w);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int c = min(n + e, cmax);
        f[e] = fmaf(gr.rstd, f[e] - gr.mu * __ldg(gr.sg + c), __ldg(gr.tb + c));
      }
    }
  }
  if (M::bias(fl)) {
    if (M::biasvec(fl, n, p.cout)) {  // 4 broadcast 16-byte loads instead of 16 clamped scalar ones
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
  if (M::rowb(fl) && row_ok) {
    const float* rb = p.rowbias + (int64_t)b * p.ld_rowbias;
    if (p.rowbias_vec && M::full(n, p.cout)) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(rb + n) + k);
        f[4 * k] += t.x; f[4 * k + 1] += t.y; f[4 * k + 2] += t.z; f[4 * k + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) f[e] += __ldg(rb + min(n + e, cmax));
    }
  }
  if (M::o2(fl) && M::o2pre(fl)) stg_store_bf16_row(stage2, lane, f);
  if (M::gelu(fl)) {
#pragma unroll
    for (int e = 0; e < 16; ++e) f[e] = gelu_exact(f[e]);
  } else if (M::silu_(fl)) {
#pragma unroll
    for (int e = 0; e < 16; ++e) f[e] = silu(f[e]);
  }
  if (M::res(fl)) {
    if (M::res32(fl)) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint4 t = lds128(stg_f32(stage, lane, k));
        f[4 * k] += __uint_as_float(t.x); f[4 * k + 1] += __uint_as_float(t.y);
        f[4 * k + 2] += __uint_as_float(t.z); f[4 * k + 3] += __uint_as_float(t.w);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const uint4 t = lds128(stg_bf16(stage, lane, k));
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
  if (M::obf16(fl)) {
#pragma unroll
    for (int e = 0; e < 16; ++e) f[e] = __bfloat162float(__float2bfloat16_rn(f[e]));
  }
  if (M::stats(fl) && row_ok) {
    if (M::full(n, p.cout)) {
#pragma unroll
      for (int e = 0; e < 16; ++e) { s1 += f[e]; s2 = fmaf(f[e], f[e], s2); }
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e)
        if (n + e <= cmax) { s1 += f[e]; s2 += f[e] * f[e]; }
    }
  }
  if (!M::obf16(fl)) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      sts128(stg_f32(stage, lane, k), __float_as_uint(f[4 * k]), __float_as_uint(f[4 * k + 1]), __float_as_uint(f[4 * k + 2]),
             __float_as_uint(f[4 * k + 3]));
  } else {
    stg_store_bf16_row(stage, lane, f);
  }
  if (M::o2(fl) && !M::o2pre(fl)) stg_store_bf16_row(stage2, lane, f);
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

struct EpiMaps {
  CUtensorMap out, res, out2;  // 32-row x kEC-column boxes of the output / residual / bf16-copy tensors
};

// Statically compiled epilogue modes (index -> flag set); every other combination runs the EM_DYN loop.  Each mode is
// its own instantiation of the CTA-pair kernel (one switch over all modes inside a single kernel pushed it past the
// 168-register budget of a 384-thread CTA: 3-5 KB of spills); the instantiations are spread over conv_pair_modes*.cu
// so that they compile in parallel.
#define SBM_EPI_MODES(X)                                                                         \
  X(0, EM_GN | EM_GELU | EM_OBF16 | EM_STATS)  /* ConvNeXt 3x3 #1: folded norm, GELU, bf16 hidden, statistics */ \
  X(1, EM_GN | EM_RES32)                       /* ConvNeXt 3x3 #2 + residual */                  \
  X(2, EM_GN | EM_RES32 | EM_O2 | EM_STATS)    /* ... + bf16 copy + statistics of the block output */ \
  X(3, EM_GN | EM_RES32 | EM_O2)                                                                  \
  X(4, EM_GN | EM_RES32 | EM_OBF16)            /* last block: bf16 only */                        \
  X(5, EM_BIAS)                                /* 1x1 res_conv, transposed conv */                \
  X(6, EM_GN)                                  /* to_qkv with the PreNorm folded */               \
  X(7, EM_BIAS | EM_STATS)                     /* attention to_out (statistics for its GroupNorm) */ \
  X(8, EM_BIAS | EM_O2)                        /* stem, down / up-sampling convs with a bf16 copy */ \
  X(9, EM_BIAS | EM_RES32)                     /* middle attention to_out + residual */           \
  X(10, EM_GN | EM_RES32 | EM_STATS)                                                             \
  X(11, EM_BIAS | EM_GELU | EM_OBF16 | EM_STATS)                                                 \
  X(12, EM_GN | EM_OBF16)                      /* to_qkv with bf16 output (16x16 / 8x8 attention) */ \
  X(13, 0u)                                    /* training: data-gradient GEMMs (plain fp32 output) */ \
  X(14, EM_BIAS | EM_GELU | EM_OBF16 | EM_STATS | EM_O2PRE)  /* training: 3x3 #1 keeps its pre-activation */ \
  X(15, EM_BIAS | EM_RES32 | EM_STATS)         /* training: 3x3 #2 + residual + statistics */     \
  X(16, EM_BIAS | EM_ROWB)                     /* UNetModel ResBlock conv #1 + time / z embedding row bias */ \
  X(17, EM_BIAS | EM_RES32 | EM_O2)            /* UNetModel ResBlock conv #2 + skip + bf16 copy */
inline int find_epi_mode(uint32_t bits) {
#define SBM_EPI_FIND(idx, mode) if (bits == (uint32_t)(mode)) return idx;
  SBM_EPI_MODES(SBM_EPI_FIND)
#undef SBM_EPI_FIND
  return -1;
}

// One output tile of the staged epilogue for one warp: its 32 rows x (nch x 16) columns.  The TMEM load of chunk c+1
// is in flight while chunk c is computed, staged and handed to TMA (two register sets, loop unrolled by two).
struct StagedTileCtx {
  uint32_t tacc;          // TMEM address of this warp's first column
  int ncol0, nch;         // first output column, 16-column chunks
  int cj, cq, ci, cb;     // box coordinates of the warp's 32 rows
  int b;                  // sample of this thread's row
  bool row_ok;
  uint8_t* wst;           // this warp's staging buffers
  uint32_t wst_u32;
  uint64_t* rbar;
  uint32_t res_bytes;
};
// Staging-buffer rotation of one epilogue warp (8 KB).  A chunk's buffer can be rewritten once the TMA store that
// read it has finished READING shared memory; that takes ~1.2k cycles under load, so with the residual layout (three
// main buffers, two chunks in flight) a warp could not stage faster than one chunk per ~600 cycles -- the pace of every
// K-short layer after the epilogue code itself had been made straight-line (pipeline trace, tools/trace_pair.py).
// Without a residual the 8 KB are cut into as many chunk buffers as fit: 4 (fp32 rows) or 8 (bf16 rows) in flight.
template <uint32_t MODE>
struct StageRing {
  using M = EpiMode<MODE>;
  static constexpr bool kDeep = !M::dyn && (MODE & (EM_RES32 | EM_RES16 | EM_O2 | EM_O2PRE)) == 0;
  static constexpr int kBytes = (MODE & EM_OBF16) ? 1024 : 2048;     // one chunk: 32 rows x 16 columns
  static constexpr int kDepth = kDeep ? (kStgPerWarp / kBytes) : 3;
};

template <uint32_t MODE>
__device__ __forceinline__ void staged_step(const ConvKernelParams& p, const EpiMaps& em, const StagedTileCtx& t, int c,
                                            const uint32_t* cur, uint32_t* nxt, int lane, uint32_t& nchunk, float& s1,
                                            float& s2, const GnRow& gr, uint32_t fl) {
  using M = EpiMode<MODE>;
  using R = StageRing<MODE>;
  const int col0 = t.ncol0 + c * kEC;
  __syncwarp();
  if constexpr (R::kDeep) {
    // buffer nchunk % kDepth was last read by the store of chunk nchunk - kDepth
    if (lane == 0) ptx::bulk_wait_group_read<R::kDepth - 1>();
    __syncwarp();
    ptx::tmem_ld_wait();                                            // `cur` has arrived
    if (c + 1 < t.nch) ptx::tmem_ld16(t.tacc + (c + 1) * kEC, nxt);  // next chunk streams in behind the math
    const uint32_t off = (nchunk % R::kDepth) * R::kBytes;
    epilogue_chunk_staged<MODE>(p, cur, col0, t.row_ok, t.b, lane, t.wst_u32 + off, 0u, s1, s2, gr, fl);
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      ptx::tma_store_5d(&em.out, t.wst + off, col0, t.cj, t.cq, t.ci, t.cb);
      ptx::bulk_commit_group();
    }
    ++nchunk;
    return;
  }
  // buffers of chunk nchunk+1 (main) / nchunk (bf16 copy) were last used by the stores of chunk nchunk-2
  if (lane == 0) {
    ptx::bulk_wait_group_read<1>();
    if (M::res(fl) && c + 1 < t.nch) {
      uint64_t* rb = &t.rbar[(nchunk + 1) % 3];
      ptx::mbar_expect_tx(rb, t.res_bytes);
      ptx::tma_load_5d(t.wst + ((nchunk + 1) % 3) * kStgMain, &em.res, rb, col0 + kEC, t.cj, 0, t.ci, t.cb);
    }
  }
  __syncwarp();
  ptx::tmem_ld_wait();                                            // `cur` has arrived
  if (c + 1 < t.nch) ptx::tmem_ld16(t.tacc + (c + 1) * kEC, nxt);  // next chunk streams in behind the math
  const uint32_t stage = t.wst_u32 + (nchunk % 3) * kStgMain;
  const uint32_t stage2 = t.wst_u32 + 3 * kStgMain + (nchunk & 1) * kStgOut2;
  if (M::res(fl)) ptx::mbar_wait(&t.rbar[nchunk % 3], (nchunk / 3) & 1);
  epilogue_chunk_staged<MODE>(p, cur, col0, t.row_ok, t.b, lane, stage, stage2, s1, s2, gr, fl);
  ptx::fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    ptx::tma_store_5d(&em.out, t.wst + (stage - t.wst_u32), col0, t.cj, t.cq, t.ci, t.cb);
    if (M::o2(fl)) ptx::tma_store_5d(&em.out2, t.wst + (stage2 - t.wst_u32), col0, t.cj, t.cq, t.ci, t.cb);
    ptx::bulk_commit_group();
  }
  ++nchunk;
}
template <uint32_t MODE>
__device__ __forceinline__ void staged_tile(const ConvKernelParams& p, const EpiMaps& em, const StagedTileCtx& t, int lane,
                                         uint32_t& nchunk, float& s1, float& s2, const GnRow& gr) {
  const uint32_t fl = EpiMode<MODE>::dyn ? epi_flags(p) : 0u;
  uint32_t va[16], vb[16];
  if (t.nch > 0) ptx::tmem_ld16(t.tacc, va);
#pragma unroll 1
  for (int c = 0; c < t.nch; c += 2) {
    staged_step<MODE>(p, em, t, c, va, vb, lane, nchunk, s1, s2, gr, fl);
    if (c + 1 < t.nch) staged_step<MODE>(p, em, t, c + 1, vb, va, lane, nchunk, s1, s2, gr, fl);
  }
}

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
  int32_t splits;   // K slices per output tile (EM_SPLITK kernels; 1 otherwise); `total` counts (tile, slice) items
  int32_t split_bstride;   // samples per workspace slab: slice s stores at sample coordinate b + s * split_bstride
};

template <int BN, int STAGES, bool kStaged, uint32_t MODE>
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
  constexpr bool kSplit = (MODE & EM_SPLITK) != 0;
  // work item -> (output tile, K slice): consecutive items are the slices of one tile
  auto split_of = [&](int& t) -> int {
    if constexpr (kSplit) {
      const int s = t % sch.splits;
      t /= sch.splits;
      return s;
    }
    return 0;
  };

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
        int t = tile_of_round(k);
        if (t >= sch.total) continue;
        const int sp = split_of(t);
        int ph, nt, b0, oh0, pj;
        tile_origin(t, ph, nt, b0, oh0, pj);
        const TapTable& tt = p.taps[ph];
        const int n0 = nt * BN + (int)rank * (BN / 2);
        SBM_TRACE(0, 0, k);   // producer reaches tile k
        auto issue = [&](int tap, int cb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if ((cb & 15) == 0) SBM_TRACE(0, 1, k);   // got a free stage (every 16th K block: the record itself costs ~0.4 us)
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
          const uint32_t lead_bar = ptx::mapa_u32(ptx::smem_u32(&full_bar[stage]), 0);
          ptx::tma_load_5d_2sm(sa, &tmA, lead_bar, cb * kBK, pj + tt.dw[tap], tt.q[tap], oh0 + tt.dh[tap], b0);
          ptx::tma_load_3d_2sm(sb, &tmB, lead_bar, cb * kBK, n0, tt.wtap[tap]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        };
        if constexpr (kSplit) {   // standard tiling only: every tap of the table is valid, K block kb = tap * cblocks + cb
          const int nkb = tt.ntaps * p.cblocks;
          const int lo = sp * nkb / sch.splits, hi = (sp + 1) * nkb / sch.splits;
          for (int kb = lo; kb < hi; ++kb) issue(kb / p.cblocks, kb % p.cblocks);
        } else {
          for (uint32_t tm = tap_mask(tt, oh0, pj); tm != 0; tm &= tm - 1) {
            const int tap = __ffs(tm) - 1;
            for (int cb = 0; cb < p.cblocks; ++cb) issue(tap, cb);
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
        int t = tile_of_round(k);
        if (t >= sch.total) continue;
        const int sp = split_of(t);
        int ph, nt, b0, oh0, pj;
        tile_origin(t, ph, nt, b0, oh0, pj);
        int num_kb = __popc(tap_mask(p.taps[ph], oh0, pj)) * p.cblocks;
        if constexpr (kSplit) num_kb = (sp + 1) * num_kb / sch.splits - sp * num_kb / sch.splits;
        SBM_TRACE(1, 0, k);   // issuer reaches tile k
        ptx::mbar_wait(&tempty_bar[astage], aphase ^ 1);
        SBM_TRACE(1, 1, k);   // accumulator stage free
        ptx::tc_fence_after_sync();
        const uint32_t tacc = tmem_base + (uint32_t)(astage * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          if ((kb & 15) == 0) SBM_TRACE(1, 2, k);   // operands of a K block landed (every 16th)
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
        SBM_TRACE(1, 3, k);   // tile committed
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
    const uint32_t wst_u32 = ptx::smem_u32(wst);
    uint64_t* rbar = res_bar + 3 * e;
    const bool has_res = p.residual != nullptr;
    const uint32_t res_bytes = 32u * kEC * (p.res_dtype == SBM_F32 ? 4u : 2u);
    uint32_t nchunk = 0;  // chunks this warp has staged so far (buffer rotation + barrier parity)
    int astage = 0;
    uint32_t aphase = 0;
    for (int k = 0; k * npairs < sch.total; ++k) {
      int t = tile_of_round(k);
      if (t >= sch.total) continue;
      const int sp = split_of(t);
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
        const int cj = pj + sub_j, ci = oh0 + sub_i, cq = tt.out_q;
        const int cb = b0 + sub_b + (kSplit ? sp * sch.split_bstride : 0);
        const int ncol0 = nt * BN + hc * (BN / 2);
        const int nch = min((BN / 2) / kEC, max(0, (p.cout - ncol0 + kEC - 1) / kEC));
        if (has_res && nch > 0 && lane == 0) {
          ptx::bulk_wait_group_read<1>();
          uint64_t* rb = &rbar[nchunk % 3];
          ptx::mbar_expect_tx(rb, res_bytes);
          ptx::tma_load_5d(wst + (nchunk % 3) * kStgMain, &em.res, rb, ncol0, cj, 0, ci, cb);
        }
        if (e == 0 && lane == 0) SBM_TRACE(2, 0, k);   // epilogue warp 0 reaches tile k
        ptx::mbar_wait(&tfull_bar[astage], aphase);
        if (e == 0 && lane == 0) SBM_TRACE(2, 1, k);   // accumulator complete
        ptx::tc_fence_after_sync();
        StagedTileCtx tc;
        tc.tacc = tmem_base + (uint32_t)(astage * BN) + (uint32_t(ew * 32) << 16) + (uint32_t)(hc * (BN / 2));
        tc.ncol0 = ncol0; tc.nch = nch;
        tc.cj = cj; tc.cq = cq; tc.ci = ci; tc.cb = cb;
        tc.b = b; tc.row_ok = row_ok;
        tc.wst = wst; tc.wst_u32 = wst_u32; tc.rbar = rbar; tc.res_bytes = res_bytes;
        staged_tile<MODE>(p, em, tc, lane, nchunk, s1, s2, gr);
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
      if (e == 0 && lane == 0) SBM_TRACE(2, 2, k);   // accumulator stage released
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

template <int BN, int STAGES, bool kStaged, uint32_t MODE>
inline int launch_conv_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiMaps& em,
                            const ConvKernelParams& p, int m_tiles, int n_tiles, int nphase, cudaStream_t stream,
                            int splits = 1, int split_bstride = 0) {
  using L = Smem2Layout<BN, STAGES, kStaged>;
  static_assert(L::kTotal <= 232448, "shared-memory budget of one CTA exceeded");
  static bool configured = false;
  if (!configured) {
    SBM_CUDA_OK(cudaFuncSetAttribute(conv_igemm_pair_kernel<BN, STAGES, kStaged, MODE>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  PairSchedule sch;
  sch.m_tiles = m_tiles;
  sch.m_pairs = (m_tiles + 1) / 2;
  sch.n_tiles = n_tiles;
  sch.nphase = nphase;
  sch.splits = splits;
  sch.split_bstride = split_bstride;
  sch.total = nphase * sch.m_pairs * n_tiles * splits;
  const int pairs = std::min(sch.total, sm_count() / 2);
  conv_igemm_pair_kernel<BN, STAGES, kStaged, MODE><<<dim3(2 * pairs), 384, L::kTotal, stream>>>(tmA, tmB, em, p, sch);
  SBM_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}


// Launch of the statically compiled epilogue mode `mode_idx` (SBM_EPI_MODES) with N tile `bn`; defined in
// conv_pair_modes*.cu.  Returns -1 when that (mode, bn) pair has no instantiation.
int launch_pair_static(int mode_idx, int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiMaps& em,
                       const ConvKernelParams& p, int m_tiles, int n_tiles, int nphase, cudaStream_t stream);

// one group of modes per translation unit
#define SBM_EPI_LAUNCH_CASE(idx, mode)                                                                                  \
  case idx:                                                                                                             \
    if (bn == 256) return launch_conv_pair<256, 5, true, (uint32_t)(mode)>(tmA, tmB, em, p, m_tiles, n_tiles, nphase, stream); \
    if (bn == 128) return launch_conv_pair<128, 6, true, (uint32_t)(mode)>(tmA, tmB, em, p, m_tiles, n_tiles, nphase, stream); \
    if (bn == 64) return launch_conv_pair<64, 8, true, (uint32_t)(mode)>(tmA, tmB, em, p, m_tiles, n_tiles, nphase, stream); \
    return -1;

}  // namespace sbm
