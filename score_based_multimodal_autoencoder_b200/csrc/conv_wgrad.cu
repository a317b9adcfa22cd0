// Weight gradient of the implicit-GEMM convolutions on the tcgen05 tensor cores.
//
//   dW[tap][o][i] = sum over output pixels p of  dY[p][o] * X[p shifted by tap][i]
//
// GEMM view: M = output channels (128 per CTA), N = input channels (128 per CTA), K = pixels, walked in blocks of
// 64 pixels.  Both operands are read exactly as they live in HBM (channels-last): a TMA box of 64 pixels x 64
// channels lands in shared memory as 64 rows (K) of 128 bytes (64 channels = the MN index), i.e. an MN-major
// operand tile; the instruction descriptor selects MN-major for A and B, so no transposed copies are ever made.
// The tap shift, zero padding and stride-2 / transposed-conv geometry reuse the forward kernel's 5-D views.
// Split-K over pixel ranges (grid.z); the partial tiles are accumulated into the packed fp32 gradient
// [tap][o][pad8(i)] by TMA reduce-add boxes (cp.reduce.async.bulk.tensor .add) staged through shared memory.
//
// Backward of: nn.Conv2d / nn.ConvTranspose2d / nn.Linear weights of unet_model.py / unet_openai.py, as needed by
// loss.backward() in train_lat_celebhq_unet_cont2.py:98-100.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>

#include "../../include/sbmae_b200.h"
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace sbm {
extern std::atomic<unsigned long long> g_launches;

constexpr int kWgBM = 128;   // output channels per CTA
constexpr int kWgPix = 64;   // pixels per K block
constexpr int kWgMaxTaps = 16;
// input channels per CTA = template parameter BN (128 or 256).  The kernel streams both operands from L2 once per
// 64-pixel block: (128 + BN) * 64 * 2 bytes for 2 * 128 * BN * 64 FLOP, i.e. 64 FLOP/B at BN = 128 and 85 FLOP/B at
// BN = 256 -- the wide tile is what keeps layers with >= 256 input channels off the L2-bandwidth floor.

struct WgTap {
  int8_t x_dh, x_dw, y_dh, y_dw;
  int16_t x_q, y_q;
  int16_t wtap;
  int16_t pad;
};
struct WgParams {
  int32_t batch, log_gw, log_gh, log_th;  // pixel grid of the GEMM (per phase) and rows per 64-pixel block
  int32_t cin, cout, cin_pad;
  int32_t num_pix_blocks, blocks_per_split;
  int32_t n_i_tiles;
  int32_t use_tma_reduce;   // 1: split-K accumulation by TMA reduce-add boxes; 0: per-thread atomicAdd
  float* dwpk;
  WgTap taps[kWgMaxTaps];
};

template <int BN>
struct WgSmem {
  static constexpr int kStages = BN == 128 ? 6 : 4;
  static constexpr int kYBytes = kWgBM * kWgPix * 2;  // atoms of 64 ch x 64 px
  static constexpr int kXBytes = BN * kWgPix * 2;
  static constexpr int kStageBytes = kYBytes + kXBytes;
  static constexpr int kBarOffset = kStages * kStageBytes;
  static constexpr int kTotal = kBarOffset + (2 * kStages + 1) * 8 + 16 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                  const __grid_constant__ CUtensorMap tmW, const __grid_constant__ WgParams p) {
  using L = WgSmem<BN>;
  constexpr int kWgBN = BN;
  constexpr int kWgStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kBarOffset);
  uint64_t* empty_bar = full_bar + kWgStages;
  uint64_t* tmem_full_bar = empty_bar + kWgStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int it = blockIdx.x % p.n_i_tiles;
  const int ot = blockIdx.x / p.n_i_tiles;
  const WgTap tp = p.taps[blockIdx.y];
  const int i0 = it * kWgBN, o0 = ot * kWgBM;
  const int kb0 = blockIdx.z * p.blocks_per_split;
  const int kb1 = min(kb0 + p.blocks_per_split, p.num_pix_blocks);
  const int log_g = p.log_gh + p.log_gw;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmY);
    ptx::prefetch_tmap(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(tmem_full_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<kWgBN>(tmem_slot);
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (kb1 > kb0) {
    if (warp == 0) {
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          int b0, h0;
          if (log_g >= 6) {
            const int per_img = 1 << (log_g - 6);
            b0 = kb / per_img;
            h0 = (kb % per_img) << p.log_th;
          } else {
            b0 = kb << (6 - log_g);
            h0 = 0;
          }
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sy = smem + stage * L::kStageBytes;
          uint8_t* sx = sy + L::kYBytes;
          ptx::mbar_expect_tx(&full_bar[stage], L::kStageBytes);
#pragma unroll
          for (int a = 0; a < kWgBM / 64; ++a)
            ptx::tma_load_5d(sy + a * (kWgPix * 128), &tmY, &full_bar[stage], o0 + a * 64, tp.y_dw, tp.y_q,
                             h0 + tp.y_dh, b0);
#pragma unroll
          for (int a = 0; a < kWgBN / 64; ++a)
            ptx::tma_load_5d(sx + a * (kWgPix * 128), &tmX, &full_bar[stage], i0 + a * 64, tp.x_dw, tp.x_q,
                             h0 + tp.x_dh, b0);
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = ptx::make_idesc_bf16(kWgBM, kWgBN, /*a_mn_major=*/1, /*b_mn_major=*/1);
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after_sync();
          const uint32_t sy = ptx::smem_u32(smem + stage * L::kStageBytes);
          const uint32_t sx = sy + L::kYBytes;
#pragma unroll
          for (int k = 0; k < kWgPix / 16; ++k) {
            // 16 pixels (K) = 16 rows of 128 B = two 8-row swizzle groups: advance by 2048 B
            const uint64_t adesc = ptx::make_desc_mn_sw128(sy + k * 2048, kWgPix * 128);
            const uint64_t bdesc = ptx::make_desc_mn_sw128(sx + k * 2048, kWgPix * 128);
            ptx::umma_bf16(tmem_base, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(tmem_full_bar);
      }
    } else if (warp >= 4) {
      // Split-K accumulation by TMA reduce-add: a warp stages its 32 rows (output channels) x 32 columns (input
      // channels) as 128-byte swizzled shared-memory rows and hands the box to cp.reduce.async.bulk.tensor (.add,
      // fp32), so the L2 sees full 128-byte row segments instead of 32 scattered 4-byte atomics per instruction (the
      // per-thread atomicAdd epilogue made the 2x2- and 4x4-level layers epilogue-bound: 84 us for 9.7 GFLOP).
      // The pipeline stages are free once tmem_full_bar fires (every MMA has read its operands): reuse them.
      const int ew = warp & 3;
      ptx::mbar_wait(tmem_full_bar, 0);
      ptx::tc_fence_after_sync();
      uint8_t* wst = smem + ew * (3 * 4096);   // 3 rotating 32 x 128 B buffers per warp (1024-byte aligned)
      if (p.use_tma_reduce) {
        uint32_t nchunk = 0;
#pragma unroll 1
        for (int c0 = 0; c0 < kWgBN; c0 += 32, ++nchunk) {
          if (i0 + c0 >= p.cin_pad) break;  // warp-uniform: nothing but padding to the right
          uint32_t v[32];
          ptx::tmem_ld16(tmem_base + (uint32_t(ew * 32) << 16) + c0, v);
          ptx::tmem_ld16(tmem_base + (uint32_t(ew * 32) << 16) + c0 + 16, v + 16);
          if (lane == 0) ptx::bulk_wait_group_read<2>();   // the buffer of chunk n was last read by the group of chunk n-3
          __syncwarp();
          ptx::tmem_ld_wait();
          uint8_t* buf = wst + (nchunk % 3) * 4096;
#pragma unroll
          for (int k = 0; k < 8; ++k)   // 16-byte chunk k of row r lives at chunk k ^ (r & 7)  (SWIZZLE_128B)
            *reinterpret_cast<uint4*>(buf + lane * 128 + ((k ^ (lane & 7)) << 4)) =
                make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_reduce_add_3d(&tmW, buf, i0 + c0, o0 + ew * 32, tp.wtap);
            ptx::bulk_commit_group();
          }
        }
        if (lane == 0) ptx::bulk_wait_group<0>();   // shared memory must outlive the reads of the last groups
      } else {
        const int o = o0 + ew * 32 + lane;
        float* dst = p.dwpk + ((int64_t)tp.wtap * p.cout + o) * p.cin_pad;
#pragma unroll 1
        for (int c0 = 0; c0 < kWgBN; c0 += 16) {
          uint32_t v[16];
          ptx::tmem_ld16(tmem_base + (uint32_t(ew * 32) << 16) + c0, v);
          ptx::tmem_ld_wait();
          if (o < p.cout) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int i = i0 + c0 + e;
              if (i < p.cin) atomicAdd(dst + i, __uint_as_float(v[e]));
            }
          }
        }
      }
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<kWgBN>(tmem_base);
}

static PFN_cuTensorMapEncodeTiled_v12000 wg_encode_fn() {
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
static int wg_ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return ((1 << l) == v) ? l : -1;
}

// 5-D (c, w, q, h, b) view of a channels-last tensor with `H x W` pixels; parity=true gives the stride-2 view
static void make_view(bool parity, int C, int H, int W, int B, int64_t ld, cuuint64_t* dim, cuuint64_t* str) {
  if (!parity) {
    dim[0] = C; dim[1] = W; dim[2] = 1; dim[3] = H; dim[4] = B;
    str[0] = ld * 2; str[1] = (cuuint64_t)W * ld * 2; str[2] = (cuuint64_t)W * ld * 2;
    str[3] = (cuuint64_t)H * W * ld * 2;
  } else {
    dim[0] = C; dim[1] = W / 2; dim[2] = W + 2; dim[3] = H / 2; dim[4] = B;
    str[0] = 2 * ld * 2; str[1] = ld * 2; str[2] = (cuuint64_t)2 * W * ld * 2;
    str[3] = (cuuint64_t)H * W * ld * 2;
  }
}

static int conv_wgrad_impl(const sbm_wgrad_args* a, cudaStream_t stream) {
  SBM_CHECK_ARG(a && a->x && a->dy && a->dwpk, "sbm_conv_wgrad: null pointer");
  SBM_CHECK_ARG(a->batch > 0 && a->cin > 0 && a->cout > 0, "sbm_conv_wgrad: bad sizes");
  const int lh = wg_ilog2(a->h), lw = wg_ilog2(a->w);
  SBM_CHECK_ARG(lh >= 0 && lw >= 0 && a->h <= 64 && a->w <= 64, "sbm_conv_wgrad: spatial extent must be 2^k <= 64");
  SBM_CHECK_ARG(a->ldx % 8 == 0 && a->lddy % 8 == 0 && a->cin_pad % 4 == 0, "sbm_conv_wgrad: bad strides");
  auto encode = wg_encode_fn();
  SBM_CHECK_ARG(encode != nullptr, "sbm_conv_wgrad: cuTensorMapEncodeTiled not available");

  WgParams p;
  memset(&p, 0, sizeof(p));
  const int kWgBN = a->cin >= 256 ? 256 : 128;  // input channels per CTA (see the note at the top)
  int gh, gw, ntaps = 0;
  cuuint64_t xdim[5], xstr[4], ydim[5], ystr[4];
  if (a->kind == SBM_CONV_S1) {
    SBM_CHECK_ARG((a->kh & 1) && (a->kw & 1) && a->kh * a->kw <= kWgMaxTaps, "sbm_conv_wgrad: bad stride-1 kernel");
    gh = a->h; gw = a->w;
    make_view(false, a->cin, a->h, a->w, a->batch, a->ldx, xdim, xstr);
    make_view(false, a->cout, gh, gw, a->batch, a->lddy, ydim, ystr);
    for (int kh = 0; kh < a->kh; ++kh)
      for (int kw = 0; kw < a->kw; ++kw) {
        const int dh = kh - a->kh / 2, dw = kw - a->kw / 2;
        if (abs(dh) >= a->h || abs(dw) >= a->w) continue;  // tap never overlaps the map: gradient stays 0
        WgTap& t = p.taps[ntaps++];
        t.x_dh = (int8_t)dh; t.x_dw = (int8_t)dw; t.x_q = 0; t.y_dh = 0; t.y_dw = 0; t.y_q = 0;
        t.wtap = (int16_t)(kh * a->kw + kw);
      }
  } else if (a->kind == SBM_CONV_S2) {
    SBM_CHECK_ARG((a->kh == 4 && a->kw == 4) || (a->kh == 3 && a->kw == 3), "sbm_conv_wgrad: bad stride-2 kernel");
    gh = a->h / 2; gw = a->w / 2;
    make_view(true, a->cin, a->h, a->w, a->batch, a->ldx, xdim, xstr);
    make_view(false, a->cout, gh, gw, a->batch, a->lddy, ydim, ystr);
    for (int kh = 0; kh < a->kh; ++kh)
      for (int kw = 0; kw < a->kw; ++kw) {
        const int rh = kh - 1, rw = kw - 1;
        const int dh = (rh < 0) ? -1 : rh / 2, par_h = (rh < 0) ? 1 : (rh & 1);
        const int dw = (rw < 0) ? -1 : rw / 2, par_w = (rw < 0) ? 1 : (rw & 1);
        if ((abs(dh) >= gh && dh != 0) || (abs(dw) >= gw && dw != 0)) continue;
        WgTap& t = p.taps[ntaps++];
        t.x_dh = (int8_t)dh; t.x_dw = (int8_t)dw; t.x_q = (int16_t)(par_h * a->w + par_w);
        t.y_dh = 0; t.y_dw = 0; t.y_q = 0;
        t.wtap = (int16_t)(kh * a->kw + kw);
      }
  } else if (a->kind == SBM_CONVT_4X4_S2) {
    gh = a->h; gw = a->w;  // pixel grid = the INPUT grid; dY is read through the stride-2 parity view
    make_view(false, a->cin, a->h, a->w, a->batch, a->ldx, xdim, xstr);
    make_view(true, a->cout, 2 * a->h, 2 * a->w, a->batch, a->lddy, ydim, ystr);
    for (int kh = 0; kh < 4; ++kh)
      for (int kw = 0; kw < 4; ++kw) {
        const int ph = 1 - (kh & 1), pw = 1 - (kw & 1);
        const int dh = (ph + 1 - kh) / 2, dw = (pw + 1 - kw) / 2;
        if (abs(dh) >= a->h || abs(dw) >= a->w) continue;
        WgTap& t = p.taps[ntaps++];
        t.x_dh = (int8_t)dh; t.x_dw = (int8_t)dw; t.x_q = 0;
        t.y_dh = 0; t.y_dw = 0; t.y_q = (int16_t)(ph * (2 * a->w) + pw);
        t.wtap = (int16_t)(kh * 4 + kw);
      }
  } else {
    SBM_CHECK_ARG(false, "sbm_conv_wgrad: unknown kind %d", a->kind);
  }
  const int log_gh = wg_ilog2(gh), log_gw = wg_ilog2(gw);
  const int log_g = log_gh + log_gw;
  const int log_th = (log_g >= 6) ? (6 - log_gw) : log_gh;
  const int nb = (log_g >= 6) ? 1 : (1 << (6 - log_g));
  const int64_t npix = (int64_t)a->batch << log_g;
  p.batch = a->batch; p.log_gw = log_gw; p.log_gh = log_gh; p.log_th = log_th;
  p.cin = a->cin; p.cout = a->cout; p.cin_pad = a->cin_pad;
  p.num_pix_blocks = (int)((npix + kWgPix - 1) / kWgPix);
  p.n_i_tiles = (a->cin + kWgBN - 1) / kWgBN;
  p.dwpk = a->dwpk;
  const int n_o_tiles = (a->cout + kWgBM - 1) / kWgBM;
  const int tiles = p.n_i_tiles * n_o_tiles * ntaps;
  int splits = std::max(1, (2 * sm_count() + tiles - 1) / tiles);
  splits = std::min(splits, p.num_pix_blocks);
  p.blocks_per_split = (p.num_pix_blocks + splits - 1) / splits;
  splits = (p.num_pix_blocks + p.blocks_per_split - 1) / p.blocks_per_split;

  CUtensorMap tmX, tmY;
  const cuuint32_t box[5] = {64u, (cuuint32_t)gw, 1u, (cuuint32_t)(1 << log_th), (cuuint32_t)nb};
  const cuuint32_t ones5[5] = {1, 1, 1, 1, 1};
  CUresult cr = encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(a->x), xdim, xstr, box, ones5,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SBM_CHECK_ARG(cr == CUDA_SUCCESS, "sbm_conv_wgrad: x tensor map encode failed (%d)", (int)cr);
  cr = encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(a->dy), ydim, ystr, box, ones5,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SBM_CHECK_ARG(cr == CUDA_SUCCESS, "sbm_conv_wgrad: dy tensor map encode failed (%d)", (int)cr);

  // packed gradient [kh*kw][cout][cin_pad] fp32 as a 3-D tensor for the reduce-add boxes (32 columns x 32 rows)
  CUtensorMap tmW;
  memset(&tmW, 0, sizeof(tmW));
  static const bool force_atomic = [] { const char* e = getenv("SBM_WGRAD_ATOMIC"); return e && atoi(e) != 0; }();
  p.use_tma_reduce = 0;
  if (!force_atomic && (reinterpret_cast<uintptr_t>(a->dwpk) & 15) == 0 && a->cin_pad % 4 == 0) {
    const cuuint64_t wdim[3] = {(cuuint64_t)a->cin_pad, (cuuint64_t)a->cout, (cuuint64_t)(a->kh * a->kw)};
    const cuuint64_t wstr[2] = {(cuuint64_t)a->cin_pad * 4, (cuuint64_t)a->cout * a->cin_pad * 4};
    const cuuint32_t wbox[3] = {32u, 32u, 1u};
    const cuuint32_t ones3[3] = {1, 1, 1};
    cr = encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, a->dwpk, wdim, wstr, wbox, ones3,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    p.use_tma_reduce = (cr == CUDA_SUCCESS) ? 1 : 0;
  }

  dim3 grid((unsigned)(p.n_i_tiles * n_o_tiles), (unsigned)ntaps, (unsigned)splits);
  if (kWgBN == 256) {
    static bool configured = false;
    if (!configured) {
      SBM_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       WgSmem<256>::kTotal));
      configured = true;
    }
    conv_wgrad_kernel<256><<<grid, 256, WgSmem<256>::kTotal, stream>>>(tmX, tmY, tmW, p);
  } else {
    static bool configured = false;
    if (!configured) {
      SBM_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       WgSmem<128>::kTotal));
      configured = true;
    }
    conv_wgrad_kernel<128><<<grid, 256, WgSmem<128>::kTotal, stream>>>(tmX, tmY, tmW, p);
  }
  SBM_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// packed fp32 gradient [taps][rows][cols_pad] -> parameter-gradient layout (element strides), dst = src (overwrite)
__global__ void unpack_wgrad_kernel(const float* __restrict__ src, float* __restrict__ dst, int taps, int rows,
                                    int cols, int cols_pad, int64_t s_tap, int64_t s_row, int64_t s_col) {
  const int64_t total = (int64_t)taps * rows * cols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    // iterate in DESTINATION order for coalesced writes when s_tap == 1 (conv weights [O][I][kh*kw])
    const int t = (int)(idx % taps);
    const int64_t rc = idx / taps;
    const int c = (int)(rc % cols);
    const int r = (int)(rc / cols);
    dst[t * s_tap + r * s_row + c * s_col] = src[((int64_t)t * rows + r) * cols_pad + c];
  }
}

// Tiled variant (both sides coalesced): packed rows are read along their column axis, the parameter layout is
// written in ITS contiguous order (taps, then whichever of row / col has the smaller stride).
__global__ void __launch_bounds__(256)
unpack_wgrad_tiled_kernel(const float* __restrict__ src, float* __restrict__ dst, int taps, int rows, int cols,
                          int cols_pad, int64_t s_tap, int64_t s_row, int64_t s_col, uint32_t taps_magic) {
  extern __shared__ float tile[];  // [32 rows][33][TP], TP = taps | 1
  const int TP = taps | 1;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int n = taps * 1024;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int cc = e & 31, rr = (e >> 5) & 31, t = e >> 10;
    const int r = r0 + rr, c = c0 + cc;
    tile[(rr * 33 + cc) * TP + t] = (r < rows && c < cols) ? src[((int64_t)t * rows + r) * cols_pad + c] : 0.f;
  }
  __syncthreads();
  const bool col_inner = llabs(s_col) <= llabs(s_row);
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int rc = taps == 1 ? e : (int)__umulhi((uint32_t)e, taps_magic);  // e / taps (exact for e < 2^16)
    const int t = e - rc * taps;
    const int inner = rc & 31, outer = rc >> 5;
    const int rr = col_inner ? outer : inner, cc = col_inner ? inner : outer;
    const int r = r0 + rr, c = c0 + cc;
    if (r < rows && c < cols) dst[t * s_tap + (int64_t)r * s_row + (int64_t)c * s_col] = tile[(rr * 33 + cc) * TP + t];
  }
}

// All weight-gradient unpacks of a backward pass (or of one gradient bucket) in ONE launch: block -> (descriptor, 32 x 32
// tile) by binary search over first_block, then the tiled body above.  99 launches of 5-20 us each (one per GEMM weight:
// 128-block grids, latency-bound at 0.3-0.7 TB/s, 1.2 ms of a CelebA training step) become one wave-filling launch.
__global__ void __launch_bounds__(256)
unpack_wgrad_multi_kernel(const sbm_pack_desc* __restrict__ descs, int n_descs) {
  extern __shared__ float tile[];  // [32 rows][33][TP]
  int lo = 0, hi = n_descs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].first_block <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const sbm_pack_desc d = descs[lo];
  const int t_idx = (int)blockIdx.x - d.first_block;
  const int taps = d.taps, rows = d.rows, cols = d.cols, cols_pad = d.cols_pad;
  const int TP = taps | 1;
  const int r0 = (t_idx / d.tiles_c) * 32, c0 = (t_idx % d.tiles_c) * 32;
  const float* src = d.src;
  float* dst = static_cast<float*>(d.dst);
  const int n = taps * 1024;
  // read side: 16-byte loads, issued in batches of up to 9 per thread BEFORE any of them is consumed (the kernel is
  // bound by memory latency: with one 4-byte load in flight per thread it moved 1.5 TB/s, ncu round 2)
  if ((cols_pad & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int n4 = taps * 256;   // (t, rr, c4): 8 float4 per 32-column row
    for (int base = 0; base < n4; base += 9 * 256) {
      float4 v[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int e = base + i * 256 + (int)threadIdx.x;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < n4) {
          const int c4 = e & 7, rr = (e >> 3) & 31, t = e >> 8;
          const int r = r0 + rr, c = c0 + c4 * 4;
          if (r < rows && c < cols_pad) v[i] = __ldcs(reinterpret_cast<const float4*>(src + ((int64_t)t * rows + r) * cols_pad + c));
        }
      }
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const int e = base + i * 256 + (int)threadIdx.x;
        if (e < n4) {
          const int c4 = e & 7, rr = (e >> 3) & 31, t = e >> 8;
          float* tp = tile + (rr * 33 + c4 * 4) * TP + t;
          tp[0] = v[i].x; tp[TP] = v[i].y; tp[2 * TP] = v[i].z; tp[3 * TP] = v[i].w;
        }
      }
    }
  } else {
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      const int cc = e & 31, rr = (e >> 5) & 31, t = e >> 10;
      const int r = r0 + rr, c = c0 + cc;
      tile[(rr * 33 + cc) * TP + t] = (r < rows && c < cols) ? src[((int64_t)t * rows + r) * cols_pad + c] : 0.f;
    }
  }
  __syncthreads();
  const bool col_inner = llabs(d.s_col) <= llabs(d.s_row);
  // write side, nn.Conv2d layout [rows][cols][taps] with odd taps: a tile row is ONE contiguous run of 32 * taps floats
  // that maps 1:1 onto the shared-memory row (TP == taps) -> 16-byte stores
  if (d.s_tap == 1 && d.s_col == taps && (taps & 1) && (d.s_row & 3) == 0 && c0 + 32 <= cols &&
      (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    const int run4 = taps * 8;
    for (int e = threadIdx.x; e < 32 * run4; e += 256) {
      const int rr = e / run4, j4 = e - rr * run4;
      const int r = r0 + rr;
      if (r < rows) {
        const float* tp = tile + rr * 33 * TP + j4 * 4;
        *reinterpret_cast<float4*>(dst + (int64_t)r * d.s_row + (int64_t)c0 * taps + j4 * 4) =
            make_float4(tp[0], tp[1], tp[2], tp[3]);
      }
    }
    return;
  }
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int rc = taps == 1 ? e : (int)__umulhi((uint32_t)e, d.taps_magic);  // e / taps (exact for e < 2^16)
    const int t = e - rc * taps;
    const int inner = rc & 31, outer = rc >> 5;
    const int rr = col_inner ? outer : inner, cc = col_inner ? inner : outer;
    const int r = r0 + rr, c = c0 + cc;
    if (r < rows && c < cols) dst[t * d.s_tap + (int64_t)r * d.s_row + (int64_t)c * d.s_col] = tile[(rr * 33 + cc) * TP + t];
  }
}

}  // namespace sbm

extern "C" {

int sbm_unpack_wgrad_multi(const sbm_pack_desc* descs_dev, int32_t n_descs, int32_t n_blocks, int32_t max_taps,
                           void* stream) {
  SBM_CHECK_ARG(descs_dev && n_descs > 0 && n_blocks > 0 && max_taps > 0 && max_taps <= 16,
                "sbm_unpack_wgrad_multi: bad args");
  const size_t smem = (size_t)(max_taps | 1) * 32 * 33 * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    SBM_CUDA_OK(cudaFuncSetAttribute(sbm::unpack_wgrad_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    configured = smem;
  }
  sbm::unpack_wgrad_multi_kernel<<<n_blocks, 256, smem, static_cast<cudaStream_t>(stream)>>>(descs_dev, n_descs);
  SBM_CUDA_OK(cudaGetLastError());
  sbm::g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

int sbm_conv_wgrad(const sbm_wgrad_args* a, void* stream) {
  return sbm::conv_wgrad_impl(a, static_cast<cudaStream_t>(stream));
}

int sbm_unpack_wgrad(const float* src, float* dst, int32_t taps, int32_t rows, int32_t cols, int32_t cols_pad,
                     int64_t s_tap, int64_t s_row, int64_t s_col, void* stream) {
  SBM_CHECK_ARG(src && dst && taps > 0 && rows > 0 && cols > 0, "sbm_unpack_wgrad: bad args");
  if ((int64_t)rows * cols >= 4096 && taps <= 16) {
    const size_t smem = (size_t)(taps | 1) * 32 * 33 * sizeof(float);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      SBM_CUDA_OK(cudaFuncSetAttribute(sbm::unpack_wgrad_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
      configured = smem;
    }
    dim3 grid((cols + 31) / 32, (rows + 31) / 32);
    sbm::unpack_wgrad_tiled_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(src, dst, taps, rows, cols,
                                                                                           cols_pad, s_tap, s_row, s_col,
                                                                                           (uint32_t)(taps > 1 ? ((1ull << 32) + taps - 1) / taps : 0));
  } else {
    const int64_t total = (int64_t)taps * rows * cols;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sbm::sm_count() * 8);
    sbm::unpack_wgrad_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, taps, rows, cols, cols_pad,
                                                                                    s_tap, s_row, s_col);
  }
  SBM_CUDA_OK(cudaGetLastError());
  sbm::g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // extern "C"
