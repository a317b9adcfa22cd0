// Depthwise 7x7 convolution on the tensor cores (unet_model.py:103-105: ds_conv + bias + time condition, plus the
// GroupNorm statistics of its output).
//
// The FFMA2 kernel in net_ops.cu does ~39 multiply-adds per output on the fp32 pipe and sits at 1.9 TB/s (0.3 of the
// HBM roofline, 39 % of the FMA peak: it cannot get within 2x of the memory bound).  Here every channel is a batch of
// small banded matrix products on bf16 mma.sync (fp32 accumulation):
//
//   out[c][To] (16 positions x 8 samples) += A[c][Ti - To] (16 x 16 Toeplitz block of the 49 taps) * x[c][Ti] (16 positions x 8 samples)
//
// positions are cut into tiles of 16 (one row of a 16x16 map, two rows of an 8x8 map, a whole 4x4 map); an output tile
// depends on the input tiles within +-3 rows.  The Toeplitz blocks (7 / 5 / 1 per channel) are built ONCE per CTA from
// the fp32 taps and live in registers as mma A fragments; a warp owns two channels of the block's 16-channel slab.
// ~2.3x the multiply-adds of the direct form, but on a pipe with two orders of magnitude more throughput: what is
// left is data movement.
//
// Shared memory holds the tile CHANNEL-PLANAR in bf16, [16 channels][8 G samples][HW positions], position pairs packed
// in one 32-bit word -- exactly the B fragment registers of m16n8k16 (k = position, n = sample), so a fragment is two
// conflict-free LDS.32 (sample stride = HW/2 + 4 words: 4 x odd; plane stride = 2 mod 8 words keeps the transposing
// stores of the load pass conflict-free too).  Pass 1 loads the fp32 channels-last activations with coalesced 16-byte
// loads, rounds to bf16 and scatters to the planes; pass 2 runs the products and writes every finished output tile IN
// PLACE over the input tile it replaces (an input tile is read exactly once, before any output of its positions is
// final); pass 3 gathers the planes back into channels-last bf16 rows with 16-byte stores.
// Operand rounding: x and the taps are rounded to bf16 (like every other convolution of the net; accumulation fp32).
// The statistics are taken from the bf16-ROUNDED outputs, like the FFMA2 kernel: they describe the tensor the folded
// GroupNorm GEMM reads.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>

#include "common.cuh"

namespace sbm {

extern std::atomic<unsigned long long> g_launches;

namespace {

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <int W>
struct DwCfg {
  static constexpr int HW = W * W;
  static constexpr int T = HW / 16;                       // position tiles of 16
  static constexpr int RPT = 16 / W;                      // map rows per tile
  static constexpr int DMAX = W == 16 ? 3 : (W == 8 ? 2 : 0);   // |Ti - To| <= DMAX
  static constexpr int ND = 2 * DMAX + 1;
  static constexpr int G = W == 16 ? 1 : (W == 8 ? 4 : 8);      // groups of 8 samples per tile
  static constexpr int NS = 8 * G;
  static constexpr int SPW = HW / 2 + 4;                  // words per sample (4 x odd: conflict-free fragments)
  static constexpr int PLW = NS * SPW + 2;                // words per channel plane (2 mod 8)
  static constexpr int RING = T < 8 ? T : 8;              // live accumulator tiles
  static constexpr int kChan = 16;
  static constexpr size_t smem = 2 * ((size_t)kChan * PLW * 4 + 8 * NS * 2 * sizeof(float));   // two tiles: planes + [warp][sample][2] sums
};

// tap of channel c that multiplies input position k of tile To + delta for output position m of tile To
template <int W>
__device__ __forceinline__ float toeplitz_tap(const float* __restrict__ w, int c, int C, int delta, int m, int k) {
  constexpr int RPT = 16 / W;
  const int mr = m / W, mc = m % W, kr = k / W, kc = k % W;
  const int dh = delta * RPT + kr - mr, dw = kc - mc;
  if (c >= C || dh < -3 || dh > 3 || dw < -3 || dw > 3) return 0.f;
  return __ldg(w + (int64_t)c * 49 + (dh + 3) * 7 + (dw + 3));
}

// 512 threads, one CTA per SM, two plane buffers.  Warps 0-7 run the products of tile j while warps 8-15 ("movers")
// write tile j-1 back to global memory and then load tile j+1 into the buffer that frees: one __syncthreads per tile,
// per-tile time = max(products, data movement) instead of their sum (the first version ran the three passes back to
// back in every CTA: 41 % / 38 % / 20 % of the samples in load / products / store, ncu source view).
template <int W>
__global__ void __launch_bounds__(512, 1)
dwconv7_mma_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                   const float* __restrict__ bias, const float* __restrict__ cond, int64_t ldc,
                   __nv_bfloat16* __restrict__ out, int64_t ldo, double* __restrict__ stats, int B, int C) {
  using K = DwCfg<W>;
  constexpr int HW = K::HW, T = K::T, DMAX = K::DMAX, ND = K::ND, G = K::G, NS = K::NS, SPW = K::SPW, PLW = K::PLW;
  constexpr int kBufWords = K::kChan * PLW;
  extern __shared__ __align__(16) uint32_t sm_all[];
  float* tstats_all = reinterpret_cast<float*>(sm_all + 2 * kBufWords);   // [2][8 warps][NS][2]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool mover = warp >= 8;
  const int mtid = tid & 255;
  const int c0 = blockIdx.x * K::kChan;
  const int ntiles = (B + NS - 1) / NS;
  const int my_tiles = blockIdx.y < ntiles ? (ntiles - 1 - (int)blockIdx.y) / (int)gridDim.y + 1 : 0;

  // ================= movers: fp32 channels-last -> bf16 channel planes (two positions per word)
  auto load_tile = [&](int j) {
    uint32_t* sm = sm_all + (j & 1) * kBufWords;
    const int s_base = ((int)blockIdx.y + j * (int)gridDim.y) * NS;
    constexpr int kItems = NS * (HW / 2) * 4;       // (sample, position pair, channel quad)
    constexpr int kLogPairs = W == 16 ? 7 : (W == 8 ? 5 : 3);
    constexpr int kBatch = 8;                       // 16 independent 16-byte loads in flight per mover thread
    static_assert(kItems % (256 * kBatch) == 0, "tile must split into whole batches");
#pragma unroll 1
    for (int i0 = mtid; i0 < kItems; i0 += 256 * kBatch) {
      float4 v0[kBatch], v1[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int i = i0 + u * 256;
        const int q = i & 3, pair = (i >> 2) & (HW / 2 - 1), n = i >> (2 + kLogPairs);
        const int b = s_base + n;
        v0[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        v1[u] = v0[u];
        if (b < B && c0 + 4 * q + 4 <= ldx) {
          const float* px = x + ((int64_t)b * HW + 2 * pair) * ldx + c0 + 4 * q;
          v0[u] = __ldg(reinterpret_cast<const float4*>(px));
          v1[u] = __ldg(reinterpret_cast<const float4*>(px + ldx));
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int i = i0 + u * 256;
        const int q = i & 3, pair = (i >> 2) & (HW / 2 - 1), n = i >> (2 + kLogPairs);
        const int cq = c0 + 4 * q;
        uint32_t* dst = sm + (4 * q) * PLW + n * SPW + pair;
        dst[0] = cq < C ? pack_bf16(v0[u].x, v1[u].x) : 0u;          // pad channels may hold anything: keep them zero
        dst[PLW] = cq + 1 < C ? pack_bf16(v0[u].y, v1[u].y) : 0u;
        dst[2 * PLW] = cq + 2 < C ? pack_bf16(v0[u].z, v1[u].z) : 0u;
        dst[3 * PLW] = cq + 3 < C ? pack_bf16(v0[u].w, v1[u].w) : 0u;
      }
    }
  };
  // ================= movers: planes -> channels-last bf16 rows (16-byte stores), statistics to global memory
  auto store_tile = [&](int j) {
    const uint32_t* sm = sm_all + (j & 1) * kBufWords;
    const float* tstats = tstats_all + (j & 1) * (8 * NS * 2);
    const int s_base = ((int)blockIdx.y + j * (int)gridDim.y) * NS;
    if (stats != nullptr && mtid < NS * 2) {
      const int b = s_base + (mtid >> 1);
      if (b < B) {
        float tsum = 0.f;
#pragma unroll
        for (int wq = 0; wq < 8; ++wq) tsum += tstats[wq * NS * 2 + mtid];   // fixed order
        atomicAdd(stats + 2 * (int64_t)b + (mtid & 1), (double)tsum);
      }
    }
    constexpr int kItems = NS * HW * 2;             // (sample, position, channel octet)
    constexpr int kLogHW = W == 16 ? 8 : (W == 8 ? 6 : 4);
    const uint16_t* hsm = reinterpret_cast<const uint16_t*>(sm);
#pragma unroll 4
    for (int i = mtid; i < kItems; i += 256) {
      const int o = i & 1, pos = (i >> 1) & (HW - 1), n = i >> (1 + kLogHW);
      const int b = s_base + n;
      if (b < B && c0 + 8 * o + 8 <= ldo) {
        const uint16_t* src = hsm + 2 * ((8 * o) * PLW + n * SPW) + pos;
        uint32_t wv[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          wv[jj] = (uint32_t)src[2 * (2 * jj) * PLW] | ((uint32_t)src[2 * (2 * jj + 1) * PLW] << 16);
        *reinterpret_cast<uint4*>(out + ((int64_t)b * HW + pos) * ldo + c0 + 8 * o) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
      }
    }
  };

  if (mover) {
    if (my_tiles > 0) load_tile(0);
    __syncthreads();
    for (int j = 0; j < my_tiles; ++j) {
      if (j >= 1) store_tile(j - 1);
      // every mover has finished reading buffer (j + 1) & 1 before anyone refills it (movers only: barrier 1)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (j + 1 < my_tiles) load_tile(j + 1);
      __syncthreads();
    }
    if (my_tiles > 0) store_tile(my_tiles - 1);
    return;
  }

  // ================= warps 0-7: banded products, outputs written in place
  const int g = lane >> 2, t = lane & 3;
  // Toeplitz blocks of this warp's two channels as mma A fragments (a0:(g,2t) a1:(g+8,2t) a2:(g,2t+8) a3:(g+8,2t+8))
  uint32_t A[2][ND][4];
  float bias_c[2];
#pragma unroll
  for (int ch = 0; ch < 2; ++ch) {
    const int c = c0 + 2 * warp + ch;
    bias_c[ch] = (bias != nullptr && c < C) ? __ldg(bias + c) : 0.f;
#pragma unroll
    for (int di = 0; di < ND; ++di) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int m = g + (r & 1) * 8, k = 2 * t + (r >> 1) * 8;
        A[ch][di][r] = pack_bf16(toeplitz_tap<W>(w, c, C, di - DMAX, m, k), toeplitz_tap<W>(w, c, C, di - DMAX, m, k + 1));
      }
    }
  }
  __syncthreads();   // tile 0 is in buffer 0
  for (int j = 0; j < my_tiles; ++j) {
    uint32_t* sm = sm_all + (j & 1) * kBufWords;
    float* tstats = tstats_all + (j & 1) * (8 * NS * 2);
    const int s_base = ((int)blockIdx.y + j * (int)gridDim.y) * NS;
#pragma unroll 1
    for (int grp = 0; grp < G; ++grp) {
      const int nb = grp * 8;
      const int bs0 = s_base + nb + 2 * t, bs1 = bs0 + 1;   // the two samples of this thread's accumulator columns
      const bool ok0 = bs0 < B, ok1 = bs1 < B;
      float s1a = 0.f, s2a = 0.f, s1b = 0.f, s2b = 0.f;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int cl = 2 * warp + ch, c = c0 + cl;
        const bool cok = c < C;
        float add0 = bias_c[ch], add1 = bias_c[ch];
        if (cond != nullptr && cok) {
          if (ok0) add0 += __ldg(cond + (int64_t)bs0 * ldc + c);
          if (ok1) add1 += __ldg(cond + (int64_t)bs1 * ldc + c);
        }
        uint32_t* plane = sm + cl * PLW;
        const uint32_t* bsrc = plane + (nb + g) * SPW + t;
        __nv_bfloat16* hplane = reinterpret_cast<__nv_bfloat16*>(plane);
        __nv_bfloat16* o0 = hplane + 2 * (nb + 2 * t) * SPW + g;    // sample 2t, position g of a tile
        __nv_bfloat16* o1 = o0 + 2 * SPW;                           // sample 2t + 1
        float acc[K::RING][4];
#pragma unroll
        for (int sl = 0; sl < K::RING; ++sl) { acc[sl][0] = 0.f; acc[sl][1] = 0.f; acc[sl][2] = 0.f; acc[sl][3] = 0.f; }

        auto emit = [&](int To, float (&a)[4]) {
          const __nv_bfloat16 h00 = __float2bfloat16_rn(a[0] + add0), h01 = __float2bfloat16_rn(a[1] + add1);
          const __nv_bfloat16 h10 = __float2bfloat16_rn(a[2] + add0), h11 = __float2bfloat16_rn(a[3] + add1);
          o0[To * 16] = h00; o0[To * 16 + 8] = h10;
          o1[To * 16] = h01; o1[To * 16 + 8] = h11;
          if (cok) {
            const float r00 = __bfloat162float(h00), r10 = __bfloat162float(h10);
            const float r01 = __bfloat162float(h01), r11 = __bfloat162float(h11);
            s1a += r00 + r10; s2a = fmaf(r00, r00, fmaf(r10, r10, s2a));
            s1b += r01 + r11; s2b = fmaf(r01, r01, fmaf(r11, r11, s2b));
          }
          a[0] = 0.f; a[1] = 0.f; a[2] = 0.f; a[3] = 0.f;   // the ring slot is reused by tile To + RING
        };

#pragma unroll
        for (int Ti = 0; Ti < T; ++Ti) {
          const uint32_t b0 = bsrc[Ti * 8], b1 = bsrc[Ti * 8 + 4];
#pragma unroll
          for (int di = 0; di < ND; ++di) {
            const int To = Ti - (di - DMAX);
            if (To >= 0 && To < T) mma16816(acc[To % K::RING], A[ch][di], b0, b1);
          }
          // every lane of the warp has read input tile Ti by now; output tile Ti - DMAX is complete and overwrites
          // input tile Ti - DMAX, which was consumed DMAX iterations ago (mma.sync is warp-synchronous)
          if (Ti - DMAX >= 0) emit(Ti - DMAX, acc[(Ti - DMAX) % K::RING]);
        }
#pragma unroll
        for (int To = (T - DMAX > 0 ? T - DMAX : 0); To < T; ++To) emit(To, acc[To % K::RING]);
      }
      if (stats != nullptr) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s1a += __shfl_xor_sync(0xffffffffu, s1a, o); s2a += __shfl_xor_sync(0xffffffffu, s2a, o);
          s1b += __shfl_xor_sync(0xffffffffu, s1b, o); s2b += __shfl_xor_sync(0xffffffffu, s2b, o);
        }
        // one slot per (warp, sample): no floating-point atomics, so the sums do not depend on the order the warps finish
        if (g == 0) {
          float* ws = tstats + (warp * NS + nb + 2 * t) * 2;
          ws[0] = s1a; ws[1] = s2a; ws[2] = s1b; ws[3] = s2b;
        }
      }
    }
    __syncthreads();   // tile j is complete (the movers store it), tile j + 1 has landed in the other buffer
  }
}

template <int W>
int launch_w(const float* x, int64_t ldx, const float* w, const float* bias, const float* cond, int64_t ldc, void* out,
             int64_t ldo, double* stats, int B, int C, cudaStream_t st) {
  using K = DwCfg<W>;
  static bool configured = false;
  static int per_sm = 1;
  if (!configured) {
    SBM_CUDA_OK(cudaFuncSetAttribute(dwconv7_mma_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::smem));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dwconv7_mma_kernel<W>, 512, K::smem) != cudaSuccess ||
        per_sm <= 0)
      per_sm = 1;
    configured = true;
  }
  const int slabs = (C + K::kChan - 1) / K::kChan;
  const int ntiles = (B + K::NS - 1) / K::NS;
  // one resident wave: every block walks several sample tiles with its Toeplitz fragments in registers
  const int gy = std::max(1, std::min(ntiles, per_sm * sm_count() / slabs));
  dwconv7_mma_kernel<W><<<dim3(slabs, gy), 512, K::smem, st>>>(x, ldx, w, bias, cond, ldc, (__nv_bfloat16*)out, ldo, stats,
                                                              B, C);
  SBM_CUDA_OK(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

}  // namespace

// Returns -1 when the tensor-core path does not apply (the caller falls back to the FFMA2 kernel).
int dwconv7_mma_launch(const float* x, int64_t ldx, const float* w, const float* bias, const float* cond, int64_t ldc,
                       void* out_bf16, int64_t ldo, double* stats, int B, int H, int W, int C, cudaStream_t st) {
  static const bool enabled = [] { const char* e = getenv("SBM_DWCONV_MMA"); return e ? atoi(e) != 0 : true; }();
  // 4x4 maps: 16 positions per sample leave the planes half padding; the FFMA2 kernel is faster there (36 vs 40 us)
  static const bool w4 = [] { const char* e = getenv("SBM_DWCONV_MMA_W4"); return e ? atoi(e) != 0 : false; }();
  if (!enabled || H != W || (W != 16 && W != 8 && !(W == 4 && w4))) return -1;
  if ((ldx & 3) || (ldo & 7) || (reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out_bf16) & 15)) return -1;
  if (W == 16) return launch_w<16>(x, ldx, w, bias, cond, ldc, out_bf16, ldo, stats, B, C, st);
  if (W == 8) return launch_w<8>(x, ldx, w, bias, cond, ldc, out_bf16, ldo, stats, B, C, st);
  return launch_w<4>(x, ldx, w, bias, cond, ldc, out_bf16, ldo, stats, B, C, st);
}

}  // namespace sbm
