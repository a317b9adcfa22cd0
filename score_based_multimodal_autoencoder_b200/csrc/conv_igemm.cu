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
#include "conv_pair.cuh"

namespace sbm {

std::atomic<unsigned long long> g_launches{0};
static bool g_force_single = false;  // debugging / A-B timing switch (sbm_conv_force_single_cta)
static thread_local int g_last_variant = 0;  // BN | pair << 16 | staged << 17 of the last launch (bench bookkeeping)
static int g_pixel_major = [] { const char* e = getenv("SBM_PIXEL_MAJOR"); return e ? atoi(e) : -1; }();       // -1: by work estimate, 0: never, 1: whenever the CTA-pair kernel runs the layer
static bool g_force_direct = false;  // A-B switch: per-thread global stores instead of the TMA-staged epilogue
#ifdef SBM_PAIR_TRACE
static unsigned long long* g_trace = nullptr;
#endif
static bool g_epi_static = [] { const char* e = getenv("SBM_EPI_STATIC"); return e ? atoi(e) != 0 : true; }();  // A-B switch: statically compiled epilogue loops (sbm_conv_epilogue_static)


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

// statically compiled epilogue modes live in conv_pair_modes{0..5}.cu (up to three modes each, compiled in parallel)
#define SBM_DECL_GROUP(g)                                                                                              \
  int launch_pair_static_g##g(int mode_idx, int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiMaps& em, \
                              const ConvKernelParams& p, int m_tiles, int n_tiles, int nphase, cudaStream_t stream);
SBM_DECL_GROUP(0) SBM_DECL_GROUP(1) SBM_DECL_GROUP(2) SBM_DECL_GROUP(3) SBM_DECL_GROUP(4) SBM_DECL_GROUP(5)
#undef SBM_DECL_GROUP
int launch_pair_static(int mode_idx, int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiMaps& em,
                       const ConvKernelParams& p, int m_tiles, int n_tiles, int nphase, cudaStream_t stream) {
  switch (mode_idx >= 16 ? 5 : mode_idx >= 12 ? 4 : mode_idx / 3) {
    case 0: return launch_pair_static_g0(mode_idx, bn, tmA, tmB, em, p, m_tiles, n_tiles, nphase, stream);
    case 1: return launch_pair_static_g1(mode_idx, bn, tmA, tmB, em, p, m_tiles, n_tiles, nphase, stream);
    case 2: return launch_pair_static_g2(mode_idx, bn, tmA, tmB, em, p, m_tiles, n_tiles, nphase, stream);
    case 3: return launch_pair_static_g3(mode_idx, bn, tmA, tmB, em, p, m_tiles, n_tiles, nphase, stream);
    case 4: return launch_pair_static_g4(mode_idx, bn, tmA, tmB, em, p, m_tiles, n_tiles, nphase, stream);
    case 5: return launch_pair_static_g5(mode_idx, bn, tmA, tmB, em, p, m_tiles, n_tiles, nphase, stream);
    default: return -1;
  }
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

// ---------------------------------------------------------------- split-K for sub-wave, K-long layers
// At 128 latents per GPU (the 8-GPU shard of the headline config) the 4x4 / 2x2 levels are GEMMs of 512..2048 rows with
// K = 4608..9216: a handful of 256 x 256 tiles.  Narrow N tiles spread them over the SMs but re-read the activation
// rows once per N tile and move 2.5 x the bytes per FLOP; an SM ingests ~40-50 B/clk, so those launches ran at 20 % of
// the tensor pipe (pipeline trace, profiles/r2b_pair_pipeline_trace_small_batch.log).  With a caller-provided fp32
// workspace the layer keeps 256-wide tiles and is cut along K instead: every (tile, K slice) work item stores its raw
// accumulators into slab `slice` of the workspace (TMA box stores), then conv_splitk_epilogue_kernel sums the slabs in
// slice order and applies the epilogue.  Deterministic (no atomics: a first version met in ONE zeroed slab through TMA
// reduce-adds, whose order made graph replays differ from the eager run in the last fp32 bits), no zeroing needed.
static bool g_splitk = [] { const char* e = getenv("SBM_SPLITK"); return e ? atoi(e) != 0 : true; }();

// number of K slices sbm_conv_igemm would use for this call if it is given a workspace (1 = no split)
static int splitk_plan(const sbm_conv_args* a) {
  if (!g_splitk || g_force_single || (a->kind != SBM_CONV_S1 && a->kind != SBM_CONV_S2) || a->out_nchw || a->cout <= 128)
    return 1;
  const int ph = a->kh / 2, pw = a->kw / 2;
  int ntaps = 0;
  int64_t M;
  if (a->kind == SBM_CONV_S1) {
    for (int kh = 0; kh < a->kh; ++kh)
      for (int kw = 0; kw < a->kw; ++kw)
        if (abs(kh - ph) < a->h && abs(kw - pw) < a->w) ++ntaps;
    M = (int64_t)a->batch * a->h * a->w;
  } else {
    // stride-2 down-sampling convolution (4x4 / 3x3, padding 1): same tap rule as the table in conv_igemm_impl.  16 taps
    // of up to 512 channels on a quarter of the pixels: at 128 latents the 8x8 -> 4x4 and 4x4 -> 2x2 layers were one
    // 128-deep K loop on 32 (or 128 narrow) CTAs, 49 us each
    const int oh = a->h / 2, ow = a->w / 2;
    for (int kh = 0; kh < a->kh; ++kh)
      for (int kw = 0; kw < a->kw; ++kw) {
        const int rh = kh - 1, rw = kw - 1;
        const int dh = (rh < 0) ? -1 : rh / 2, dw = (rw < 0) ? -1 : rw / 2;
        if ((abs(dh) >= oh && dh != 0) || (abs(dw) >= ow && dw != 0)) continue;
        ++ntaps;
      }
    M = (int64_t)a->batch * oh * ow;
  }
  const int num_kb = ntaps * ((a->cin + kBK - 1) / kBK);
  if (M < 128 || num_kb < 32) return 1;
  const int64_t tiles = ((M + 255) / 256) * ((a->cout + 255) / 256);
  const int npairs = sm_count() / 2;
  if (tiles * 2 > npairs) return 1;
  const int splits = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(8, npairs / tiles), num_kb / 8));
  // split-K runs the standard tiling, which also multiplies the taps that only read zero padding.  Where the
  // pixel-major tiling would skip them (conv_igemm_impl: 56 % of a 3x3 at 2x2, 31 % at 4x4, batch in 256-sample blocks),
  // the split has to win that work back: at 1024 latents the 2x2 level is 32 tiles, 2 slices of 9 taps lose to 1 slice
  // of 4 (measured: forward 9.16 -> 9.29 ms with the split on)
  int64_t valid = 0;
  for (int i = 0; i < a->h; ++i)
    for (int j = 0; j < a->w; ++j)
      for (int kh = 0; kh < a->kh; ++kh)
        for (int kw = 0; kw < a->kw; ++kw)
          valid += (i + kh - ph >= 0 && i + kh - ph < a->h && j + kw - pw >= 0 && j + kw - pw < a->w) ? 1 : 0;
  const double work_pm = (double)((a->batch + 255) / 256) * 256 * valid;
  const double work_std = (double)((M + kBM - 1) / kBM) * kBM * ntaps;
  if (a->kind == SBM_CONV_S1 && g_pixel_major != 0 && a->h * a->w <= 256 && work_pm < 0.97 * work_std &&
      splits * work_pm < 1.25 * work_std)
    return 1;
  return splits;
}

// samples per workspace slab: the rows of whole 256-row tile pairs (a tail tile stores its padding rows inside its own slab)
static int64_t splitk_slab_samples(int64_t M, int64_t ohw) { return ((M + 255) / 256 * 256 + ohw - 1) / ohw; }

// epilogue of a split-K convolution: thread = (output row, 16-column chunk), chunk index fastest, so a warp reads and
// writes contiguous pieces of one row (or of a few consecutive rows when the row has fewer than 32 chunks).  The K
// slices are summed in slice order: the result does not depend on which SM pair finished first.
__global__ void __launch_bounds__(256)
conv_splitk_epilogue_kernel(const __grid_constant__ ConvKernelParams p, const float* __restrict__ ws, int64_t ldw,
                            int chunks, int64_t rows, int splits, int64_t slab) {
  const int lane = threadIdx.x & 31;
  const int64_t idx = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32 + lane;
  const int64_t r = idx / chunks;
  const int chunk = (int)(idx - r * chunks);
  const int log_ohw = p.log_oh + p.log_ow;
  const bool row_ok = r < rows;
  const int64_t rr = row_ok ? r : 0;
  const int b = (int)(rr >> log_ohw);
  const int rem = (int)(rr & ((1 << log_ohw) - 1));
  const int oh = rem >> p.log_ow, j = rem & ((1 << p.log_ow) - 1);
  const int n = chunk * 16;
  float acc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) acc[e] = 0.f;
  if (row_ok) {
    const float* src = ws + rr * ldw + n;
    for (int s = 0; s < splits; ++s, src += slab) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (n + 4 * k + 4 <= ldw) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(src) + k);
          acc[4 * k] += t.x; acc[4 * k + 1] += t.y; acc[4 * k + 2] += t.z; acc[4 * k + 3] += t.w;
        }
    }
  }
  uint32_t v[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(acc[e]);
  const int64_t o_base = (int64_t)b * p.o_sb + (int64_t)oh * p.o_sh + (int64_t)j * p.o_sw;
  const int64_t r_base = (int64_t)b * p.r_sb + (int64_t)oh * p.r_sh + (int64_t)j * p.r_sw;
  const int64_t o2_base = (int64_t)b * p.o2_sb + (int64_t)oh * p.o2_sh + (int64_t)j * p.o2_sw;
  float s1 = 0.f, s2 = 0.f;
  const GnRow gr = gn_row(p, b, oh, j, row_ok);
  epilogue16(p, v, n, row_ok, b, o_base, r_base, o2_base, s1, s2, gr);
  if (p.stats != nullptr) {
    if (!row_ok) { s1 = 0.f; s2 = 0.f; }
    // a warp covers 32 / chunks consecutive rows: almost always one sample
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    if (__all_sync(0xffffffffu, b == b0 || !row_ok)) {
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0 && __shfl_sync(0xffffffffu, (int)row_ok, 0)) {
        atomicAdd(p.stats + 2 * (int64_t)b0, (double)s1);
        atomicAdd(p.stats + 2 * (int64_t)b0 + 1, (double)s2);
      }
    } else if (row_ok) {
      atomicAdd(p.stats + 2 * (int64_t)b, (double)s1);
      atomicAdd(p.stats + 2 * (int64_t)b + 1, (double)s2);
    }
  }
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
  p.rowbias_vec = (a->rowbias != nullptr && (reinterpret_cast<uintptr_t>(a->rowbias) & 15) == 0 && a->ld_rowbias % 4 == 0) ? 1 : 0;
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
  bool use_pair = pm ? true : choose(m_tiles);
  // split-K (see splitk_plan): standard tiling, 256-wide tiles, raw accumulators into the caller's zeroed workspace
  const int splits = (a->splitk_ws != nullptr && a->ld_ws >= a->cout && a->ld_ws % 4 == 0 &&
                      (reinterpret_cast<uintptr_t>(a->splitk_ws) & 15) == 0 && M >= 128) ? splitk_plan(a) : 1;
  const int64_t slab_samples = splitk_slab_samples(M, (int64_t)1 << log_ohw);
  if (splits > 1) {
    SBM_CHECK_ARG(a->ws_elems >= (int64_t)splits * slab_samples * ((int64_t)1 << log_ohw) * a->ld_ws,
                  "sbm_conv_igemm: split-K workspace of %lld floats is too small (sbm_conv_splitk_ws_elems)",
                  (long long)a->ws_elems);
    pm = false;
    m_tiles = (int)((M + kBM - 1) / kBM);
    BN = 256;
    use_pair = true;
  }
  const int n_tiles_pair = (a->cout + BN - 1) / BN;
  p.pm = pm ? 1 : 0;
#ifdef SBM_PAIR_TRACE
  p.trace = g_trace;
#endif

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

  if (splits > 1) {
    EpiMaps em;
    // workspace = [splits][slab_samples][oh][ow][ld_ws]: one tensor map, slice s stores at sample coordinate b + s * slab
    SBM_CHECK_ARG(encode_rowbox_map(encode, &em.out, a->splitk_ws, SBM_F32, a->ld_ws, a->cout, ow, oh,
                                    (int)(splits * slab_samples), 1, log_ow, log_th, false),
                  "sbm_conv_igemm: split-K workspace tensor map encode failed");
    em.res = em.out;
    em.out2 = em.out;
    ConvKernelParams q = p;   // the GEMM pass has no epilogue arithmetic
    q.bias = nullptr; q.residual = nullptr; q.out = a->splitk_ws; q.out2 = nullptr; q.stats = nullptr; q.rowbias = nullptr;
    q.gn_stats = nullptr; q.gn_tab = nullptr; q.act = SBM_ACT_NONE; q.out_dtype = SBM_F32; q.out2_preact = 0;
    q.bias_vec = 0; q.rowbias_vec = 0;
    q.taps[0].out_q = 0;
    g_last_variant = BN | (1 << 16) | (1 << 17) | (1 << 20) | (splits << 24);
    const int rc = launch_conv_pair<256, 5, true, EM_SPLITK>(tmA, tmB, em, q, m_tiles, n_tiles_pair, nphase, stream, splits,
                                                             (int)slab_samples);
    if (rc != 0) return rc;
    const int chunks = (a->cout + 15) / 16;
    const int64_t items = M * chunks;
    conv_splitk_epilogue_kernel<<<(unsigned)((items + 255) / 256), 256, 0, stream>>>(
        p, a->splitk_ws, a->ld_ws, chunks, M, splits, slab_samples * ((int64_t)1 << log_ohw) * a->ld_ws);
    SBM_CUDA_OK(cudaGetLastError());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
  }
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
      // statically compiled epilogue loop when the flag set is one of SBM_EPI_MODES
      int epi_mode = -1;
      const bool al16 = (a->bias == nullptr || p.bias_vec) &&
                        (a->gn_tab == nullptr || (reinterpret_cast<uintptr_t>(a->gn_tab) & 15) == 0);
      if (g_epi_static && (a->gn_tab == nullptr || a->cout % 4 == 0) && al16 &&
          (a->act == SBM_ACT_NONE || a->act == SBM_ACT_GELU)) {
        uint32_t bits = 0;
        if (a->gn_tab) bits |= EM_GN;
        if (a->bias) bits |= EM_BIAS;
        if (a->rowbias) bits |= EM_ROWB;
        if (a->act == SBM_ACT_GELU) bits |= EM_GELU;
        if (a->residual) bits |= (a->res_dtype == SBM_F32) ? EM_RES32 : EM_RES16;
        if (a->out_dtype == SBM_BF16) bits |= EM_OBF16;
        if (a->out2) bits |= a->out2_preact ? EM_O2PRE : EM_O2;
        if (a->stats) bits |= EM_STATS;
        epi_mode = find_epi_mode(bits);
      }
      // (BN = 64, the PolyMNIST net's 42- / 64-channel layers: K-short and all epilogue -- at 32k latents they ran the
      // EM_DYN loop at 2.5 TB/s, a fifth of a forward)
      if (epi_mode >= 0 && BN >= 64) {
        const int rc = launch_pair_static(epi_mode, BN, tmA, tmB, em, p, m_tiles, n_tiles_pair, nphase, stream);
        if (rc >= 0) {
          g_last_variant |= 1 << 19;
          return rc;
        }
      }
      if (BN == 256) return launch_conv_pair<256, 5, true, EM_DYN>(tmA, tmB, em, p, m_tiles, n_tiles_pair, nphase, stream);
      if (BN == 128) return launch_conv_pair<128, 6, true, EM_DYN>(tmA, tmB, em, p, m_tiles, n_tiles_pair, nphase, stream);
      return launch_conv_pair<64, 8, true, EM_DYN>(tmA, tmB, em, p, m_tiles, n_tiles_pair, nphase, stream);
    }
    memset(&em, 0, sizeof(em));
    if (BN == 256) return launch_conv_pair<256, 6, false, EM_DYN>(tmA, tmB, em, p, m_tiles, n_tiles_pair, nphase, stream);
    if (BN == 128) return launch_conv_pair<128, 8, false, EM_DYN>(tmA, tmB, em, p, m_tiles, n_tiles_pair, nphase, stream);
    return launch_conv_pair<64, 10, false, EM_DYN>(tmA, tmB, em, p, m_tiles, n_tiles_pair, nphase, stream);
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

#ifdef SBM_PAIR_TRACE
int sbm_debug_pair_trace(unsigned long long* buf) {
  sbm::g_trace = buf;
  return 0;
}
#endif

int sbm_conv_epilogue_static(int32_t on) {
  sbm::g_epi_static = on != 0;
  return 0;
}

int sbm_conv_splitk_plan(const sbm_conv_args* a) { return a ? sbm::splitk_plan(a) : 1; }
int64_t sbm_conv_splitk_ws_elems(const sbm_conv_args* a) {
  if (a == nullptr) return 0;
  const int splits = sbm::splitk_plan(a);
  if (splits <= 1) return 0;
  const int64_t ohw = (a->kind == SBM_CONV_S2) ? (int64_t)(a->h / 2) * (a->w / 2) : (int64_t)a->h * a->w;
  return (int64_t)splits * sbm::splitk_slab_samples((int64_t)a->batch * ohw, ohw) * ohw * a->ld_ws;
}
int sbm_conv_splitk(int32_t on) {
  sbm::g_splitk = on != 0;
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
