// Fused reverse-SDE predictor-corrector sampler steps and DSM-loss kernels (HBM-bound, fp32 state).
//
//   predictor step   : sde_helper2.py:45-52 + 277-317 + 352-356/402-407/445-450   (12 B / latent element)
//   corrector step   : sde_helper2.py:54-101  -> norms kernel (4 B/elt) + update kernel (12 B/elt)
//   observed-latent imputation : train_lat_celebhq_unet_cont2.py:293-303, 309-311 (fused as an epilogue)
//   DSM perturb/loss : sde_helper2.py:167-185
//
// Noise: either an injected buffer (parity tests feed the oracle's noise) or Philox4x32-10 keyed by
// (seed, draw id) with counter = GLOBAL element index / 4, so a batch shard on any GPU draws the same
// numbers as the unsharded batch would.  Latent tensors are dense [B, M, D, D] fp32; E = M*D*D per sample.
#include <atomic>

#include "../../include/sbmae_b200.h"
#include "common.cuh"

namespace sbm {
extern std::atomic<unsigned long long> g_launches;
static inline void count_launch_s() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ------------------------------------------------------------------------------ Philox4x32-10
struct Philox {
  static constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
  __device__ __forceinline__ static uint4 round10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(kM0, c.x), lo0 = kM0 * c.x;
      const uint32_t hi1 = __umulhi(kM1, c.z), lo1 = kM1 * c.z;
      c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
      k.x += kW0;
      k.y += kW1;
    }
    return c;
  }
};
// Box-Muller on the hardware special-function unit: u in (0,1) from the top of the 32-bit word, r = sqrt(-2 ln u)
// via MUFU.LG2 / MUFU.SQRT, angle via MUFU.SIN / MUFU.COS.  (The in-kernel stream is this library's own; the
// reference's torch-CUDA stream is reproduced by rng="torch", which injects torch.randn_like draws.)
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float2 box_muller(uint32_t x, uint32_t y) {
  const float u = x * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
  const float v = y * (2.3283064e-10f * 6.2831855f) + (2.3283064e-10f * 6.2831855f / 2.0f);
  const float s = fast_sqrt(-1.3862943611198906f * __log2f(u));  // -2 ln u = -2 ln2 log2 u
  float sn, cs;
  __sincosf(v, &sn, &cs);
  return make_float2(s * sn, s * cs);
}
// 4 standard normals for global element quad `quad` of draw `draw` under `seed`
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t draw, uint64_t quad) {
  const uint4 ctr = make_uint4((uint32_t)quad, (uint32_t)(quad >> 32), (uint32_t)draw, (uint32_t)(draw >> 32));
  const uint4 r = Philox::round10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float2 a = box_muller(r.x, r.y), b = box_muller(r.z, r.w);
  return make_float4(a.x, a.y, b.x, b.y);
}
// z0^2 + z1^2 + z2^2 + z3^2 of philox_normal4(seed, draw, quad) without forming the normals: a Box-Muller pair is
// (s sin v, s cos v) with s^2 = -2 ln u, so its sum of squares is s^2 (sin^2 + cos^2 = 1; equal to the sum over the
// materialised values up to fp32 rounding).  Used by the corrector's noise-norm reduction.
__device__ __forceinline__ float philox_sumsq4(uint64_t seed, uint64_t draw, uint64_t quad) {
  const uint4 ctr = make_uint4((uint32_t)quad, (uint32_t)(quad >> 32), (uint32_t)draw, (uint32_t)(draw >> 32));
  const uint4 r = Philox::round10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float ua = r.x * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
  const float ub = r.z * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
  return -1.3862943611198906f * (__log2f(ua) + __log2f(ub));
}
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t draw, uint64_t idx) {
  const uint64_t quad = idx >> 2;
  const uint4 ctr = make_uint4((uint32_t)quad, (uint32_t)(quad >> 32), (uint32_t)draw, (uint32_t)(draw >> 32));
  const uint4 r = Philox::round10(ctr, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t w = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
  // torch.rand convention: uniform in [0,1) from the top 24 bits
  return (float)(w >> 8) * (1.0f / 16777216.0f);
}

// n / d for any 32-bit n by multiply-high (round-up method): the elementwise kernels recover (sample, offset in
// sample) from the flat quad index without an integer division.
struct FastDiv {
  uint32_t d, m, s1, s2;
};
static FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  f.s1 = l < 1 ? l : 1;
  f.s2 = l < 1 ? 0 : l - 1;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  const uint32_t t = __umulhi(f.m, n);
  return (t + ((n - t) >> f.s1)) >> f.s2;
}

// ------------------------------------------------------------------------------ SDE scalar functions
struct SdeP {
  int kind;     // SBM_SDE_VP / SUBVP / VE
  float b0, b1; // beta_min,beta_max  or  sigma_min,sigma_max
  int N;
};
__device__ __forceinline__ void sde_drift_diff(const SdeP& s, float t, float& drift_coef, float& g) {
  if (s.kind == SBM_SDE_VE) {
    const float sigma = s.b0 * powf(s.b1 / s.b0, t);
    drift_coef = 0.f;
    g = sigma * sqrtf(2.f * (logf(s.b1) - logf(s.b0)));
  } else {
    const float beta = s.b0 + t * (s.b1 - s.b0);
    drift_coef = -0.5f * beta;
    if (s.kind == SBM_SDE_VP) {
      g = fast_sqrt(beta);  // MUFU.SQRT: 1 ulp, far inside the 1e-5 parity tolerance of the step kernels
    } else {
      const float discount = 1.f - expf(-2.f * s.b0 * t - (s.b1 - s.b0) * t * t);
      g = sqrtf(beta * discount);
    }
  }
}
__device__ __forceinline__ void sde_marginal(const SdeP& s, float t, float& mean_coef, float& std) {
  if (s.kind == SBM_SDE_VE) {
    mean_coef = 1.f;
    std = s.b0 * powf(s.b1 / s.b0, t);
  } else {
    const float lmc = -0.25f * t * t * (s.b1 - s.b0) - 0.5f * t * s.b0;
    mean_coef = expf(lmc);
    const float v = 1.f - expf(2.f * lmc);
    std = s.kind == SBM_SDE_VP ? sqrtf(v) : v;  // subVP: no sqrt (sde_helper2.py:412)
  }
}

// sde.discretize(x, t) -> f = fc-rule applied to x, G.  VP: DDPM rule from the discrete_betas table
// (sde_helper2.py:373-381: f = sqrt(alpha_i) x - x, G = sqrt(beta_i), alpha_i = 1 - beta_i in fp32 like the table the
// reference builds, :337-338); VE: SMLD rule from the discrete_sigmas table (:465-473: f = 0,
// G = sqrt(sigma_i^2 - sigma_{i-1}^2), sigma_{-1} = 0); subVP has no override and takes the Euler-Maruyama rule of the
// base class (:236-253: f = drift / N, G = diffusion * sqrt(1/N)).  i = (t (N-1) / T).long().
__device__ __forceinline__ void rd_discretize(const SdeP& s, float t, float T, const float* __restrict__ table,
                                              float& fc, float& G) {
  if (s.kind == SBM_SDE_SUBVP) {
    float dc, g;
    sde_drift_diff(s, t, dc, g);
    const float dtp = 1.f / (float)s.N;
    fc = dc * dtp;
    G = g * sqrtf(dtp);
    return;
  }
  const int i = min(max((int)(t * (float)(s.N - 1) / T), 0), s.N - 1);
  if (s.kind == SBM_SDE_VP) {
    const float beta = __ldg(table + i);
    fc = sqrtf(1.f - beta);
    G = sqrtf(beta);
  } else {
    const float sigma = __ldg(table + i);
    const float adj = i == 0 ? 0.f : __ldg(table + i - 1);
    fc = 0.f;
    G = sqrtf(sigma * sigma - adj * adj);
  }
}
__device__ __forceinline__ float rd_f(const SdeP& s, float fc, float x) {
  return s.kind == SBM_SDE_VP ? __fsub_rn(__fmul_rn(fc, x), x) : fc * x;
}

struct Impute {
  const float* z_obs;   // clean latents [B,M,D,D] or NULL (no imputation)
  uint32_t mask;        // bit m set = modality channel m observed
  int noise_obs;
  float t_next;         // time of the step the written state is the input of
  const float* t_next_dev;  // optional device scalar overriding t_next
  FastDiv ddq;          // D*D/4 quads per modality channel
};
// eq = quad index inside the sample; the 4 elements of a quad share the channel (dd % 4 == 0)
__device__ __forceinline__ float4 apply_impute(const Impute& im, float2 cf, float4 v, uint32_t q, uint32_t eq) {
  if (im.z_obs == nullptr) return v;
  const uint32_t m = fdiv(eq, im.ddq);
  if (!((im.mask >> m) & 1u)) return v;
  const float4 z = __ldg(reinterpret_cast<const float4*>(im.z_obs) + q);
  // noised observation = mean + std * z_obs with mean = exp(lmc) * z_obs (train_lat_celebhq_unet_cont2.py:296-297)
  return make_float4(__fmaf_rn(cf.y, z.x, cf.x * z.x), __fmaf_rn(cf.y, z.y, cf.x * z.y), __fmaf_rn(cf.y, z.z, cf.x * z.z),
                     __fmaf_rn(cf.y, z.w, cf.x * z.w));
}
// (mean coefficient, std) of the re-noised observation; (1, 0) = clean latent
__device__ __forceinline__ float2 impute_coef(const Impute& im, const SdeP& s) {
  if (im.z_obs == nullptr || !im.noise_obs) return make_float2(1.f, 0.f);
  float mc, sd;
  sde_marginal(s, im.t_next_dev ? __ldg(im.t_next_dev) : im.t_next, mc, sd);
  return make_float2(mc, sd);
}

// nn.Dropout(p) in training mode (unet_openai.py:265), in place on a channels-last activation x[rows][C] (fp32 or
// bf16, row stride ld): x *= keep / (1 - p), keep = (uniform >= p).  The uniform of element (row, c) is word (c & 3) of
// Philox(seed, draw, quad = (row * C8 + c) / 4) with C8 = C rounded up to 8, so the SAME call on the gradient of the
// output is the backward (the mask is regenerated, never stored).  One channel octet per thread.
template <typename T>
__global__ void __launch_bounds__(256)
dropout_kernel(T* __restrict__ x, int64_t ld, int64_t n_oct, FastDiv c8d, int C, float p, float inv_keep,
               uint64_t seed, uint64_t draw, const uint64_t* draw_dev) {
  if (draw_dev) draw += *draw_dev;
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t thr = (uint32_t)(p * 16777216.0f);  // keep iff (word >> 8) >= p * 2^24  (torch.rand's 24-bit uniform)
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < n_oct; o += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t row = fdiv((uint32_t)o, c8d);
    const int c0 = ((uint32_t)o - row * c8d.d) * 8;
    const uint64_t quad = (uint64_t)o * 2;
    const uint4 r0 = Philox::round10(make_uint4((uint32_t)quad, (uint32_t)(quad >> 32), (uint32_t)draw,
                                                (uint32_t)(draw >> 32)), key);
    const uint4 r1 = Philox::round10(make_uint4((uint32_t)(quad + 1), (uint32_t)((quad + 1) >> 32), (uint32_t)draw,
                                                (uint32_t)(draw >> 32)), key);
    const uint32_t w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    T* px = x + (int64_t)row * ld + c0;
    if constexpr (sizeof(T) == 2) {
      uint4 v = *reinterpret_cast<uint4*>(px);
      __nv_bfloat16* e = reinterpret_cast<__nv_bfloat16*>(&v);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        e[j] = (c0 + j < C && (w[j] >> 8) >= thr) ? __float2bfloat16(__bfloat162float(e[j]) * inv_keep)
                                                  : __float2bfloat16(0.f);
      *reinterpret_cast<uint4*>(px) = v;
    } else {
      float4 a = *reinterpret_cast<float4*>(px), b = *reinterpret_cast<float4*>(px + 4);
      float* e[2] = {reinterpret_cast<float*>(&a), reinterpret_cast<float*>(&b)};
#pragma unroll
      for (int j = 0; j < 8; ++j)
        e[j >> 2][j & 3] = (c0 + j < C && (w[j] >> 8) >= thr) ? e[j >> 2][j & 3] * inv_keep : 0.f;
      *reinterpret_cast<float4*>(px) = a;
      *reinterpret_cast<float4*>(px + 4) = b;
    }
  }
}


// ------------------------------------------------------------------------------ predictor
// One thread = one quad (4 consecutive latent elements), U quads in flight per loop trip.  The Philox rounds and the
// Box-Muller transform are one long dependent chain per quad: with U = 2 the kernel issued on 58 % of its cycles at 55 %
// of DRAM bandwidth (ncu, round 2) -- neither bound, waiting on its own arithmetic latency; more independent quads
// per thread fill those slots.
template <int U>
__global__ void __launch_bounds__(256)
predictor_kernel(const float4* __restrict__ x, const float4* __restrict__ score, const float* __restrict__ t,
                 const float4* __restrict__ noise, float4* __restrict__ x_out, float4* __restrict__ x_mean_out,
                 uint32_t n_quads, FastDiv eqd, SdeP s, int ode, uint64_t seed, uint64_t draw,
                 const uint64_t* draw_dev, uint64_t quad_offset, Impute im, int rd, const float* __restrict__ table,
                 float T) {
  if (draw_dev) draw += *draw_dev;
  const float dt = -1.f / (float)s.N;
  const float sq = sqrtf(-dt);
  const float2 icoef = impute_coef(im, s);
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t q0 = blockIdx.x * blockDim.x + threadIdx.x; q0 < n_quads; q0 += U * stride) {
    uint32_t qs[U];
    float4 xv[U], sv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      qs[u] = q0 + u * stride;
      if (qs[u] < n_quads) {  // read-once / write-once streams: keep them out of the way of L2-resident data
        xv[u] = __ldcs(x + qs[u]);
        sv[u] = __ldcs(score + qs[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t q = qs[u];
      if (q >= n_quads) break;
      const uint32_t b = fdiv(q, eqd);
      float4 mean;
      float gs;  // coefficient of the noise
      if (!rd) {
        float dc, g;
        sde_drift_diff(s, __ldg(t + b), dc, g);
        const float g2 = g * g * (ode ? 0.5f : 1.f);
        mean.x = xv[u].x + (dc * xv[u].x - g2 * sv[u].x) * dt;
        mean.y = xv[u].y + (dc * xv[u].y - g2 * sv[u].y) * dt;
        mean.z = xv[u].z + (dc * xv[u].z - g2 * sv[u].z) * dt;
        mean.w = xv[u].w + (dc * xv[u].w - g2 * sv[u].w) * dt;
        gs = g * sq;
      } else {
        // reverse-diffusion (ancestral) rule: (f, G) = sde.discretize(x, t); rev_f = f - G^2 s [*0.5];
        // x_mean = x - rev_f; x' = x_mean + G z   (sde_helper2.py:236-253, 319-324, 373-381, 465-473)
        float fc, G;  // f = fc * x
        rd_discretize(s, __ldg(t + b), T, table, fc, G);
        const float G2 = G * G * (ode ? 0.5f : 1.f);
        // f is formed the way the reference does (sqrt(alpha) * x - x: two roundings), then x - (f - G^2 s)
        mean.x = xv[u].x - (rd_f(s, fc, xv[u].x) - G2 * sv[u].x);
        mean.y = xv[u].y - (rd_f(s, fc, xv[u].y) - G2 * sv[u].y);
        mean.z = xv[u].z - (rd_f(s, fc, xv[u].z) - G2 * sv[u].z);
        mean.w = xv[u].w - (rd_f(s, fc, xv[u].w) - G2 * sv[u].w);
        gs = G;
      }
      float4 out = mean;
      if (!ode) {
        const float4 z = noise ? noise[q] : philox_normal4(seed, draw, quad_offset + (uint64_t)q);
        out.x += gs * z.x; out.y += gs * z.y; out.z += gs * z.z; out.w += gs * z.w;
      }
      if (x_mean_out) __stcs(x_mean_out + q, mean);
      __stcs(x_out + q, apply_impute(im, icoef, out, q, q - b * eqd.d));
    }
  }
}

// ------------------------------------------------------------------------------ corrector
// norms: acc[0] += sum_b ||grad_b||, acc[1] += sum_b ||noise_b||  (fp64 accumulators).  One warp reduces S consecutive
// samples at a time, S chosen so that S * (E/4) quads are a multiple of 32: every lane is busy on every trip (a
// PolyMNIST latent has 80 quads, i.e. 2.5 warp trips per sample).
// GRAD: read the score and accumulate acc[0].  NM: noise source of acc[1]: 0 none, 1 injected buffer, 2 Philox.
// The Philox noise norm does not depend on any data (it is a function of seed, draw id and element index), and
// regenerating the stream is pure ALU work: <true, 2> made this kernel issue-bound at 0.35 of the HBM roofline (ncu,
// round 1).  The samplers therefore run <false, 2> (no memory traffic at all) on a side stream next to the score-net
// forward whose output <true, 0> then reduces at memory speed.
template <int S, bool GRAD, int NM, int NU>
__global__ void __launch_bounds__(256)
corrector_norms_kernel(const float4* __restrict__ grad, const float4* __restrict__ noise, double* __restrict__ acc,
                       int B, int EQ, uint64_t seed, uint64_t draw, const uint64_t* draw_dev,
                       uint64_t quad_offset) {
  if (NM == 2 && draw_dev) draw += *draw_dev;
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int groups = (B + S - 1) / S;
  double a0 = 0.0, a1 = 0.0;
  for (int gidx = blockIdx.x * warps_per_block + (threadIdx.x >> 5); gidx < groups; gidx += gridDim.x * warps_per_block) {
    const int b0 = gidx * S;
    const int ns = min(S, B - b0);
    const uint32_t base = (uint32_t)b0 * (uint32_t)EQ;
    const int total = ns * EQ;
    float sg[S], sn[S];
#pragma unroll
    for (int k = 0; k < S; ++k) { sg[k] = 0.f; sn[k] = 0.f; }
    for (int q = lane; q < total; q += 32 * NU) {
      float4 gv[NU];
      float nv[NU];
#pragma unroll
      for (int u = 0; u < NU; ++u) {  // NU 16-byte loads (or Philox chains) in flight per lane
        const int qq = q + 32 * u;
        const bool on = qq < total;
        gv[u] = (GRAD && on) ? __ldcs(grad + base + qq) : make_float4(0.f, 0.f, 0.f, 0.f);
        nv[u] = 0.f;
        if (NM == 1 && on) {
          const float4 z = __ldcs(noise + base + qq);
          nv[u] = z.x * z.x + z.y * z.y + z.z * z.z + z.w * z.w;
        }
        if (NM == 2 && on) nv[u] = philox_sumsq4(seed, draw, quad_offset + (uint64_t)(base + qq));
      }
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int qq = q + 32 * u;
        const float gq = gv[u].x * gv[u].x + gv[u].y * gv[u].y + gv[u].z * gv[u].z + gv[u].w * gv[u].w;
        if (S == 1) {
          sg[0] += gq;
          sn[0] += nv[u];
        } else {
          // sample index inside the group by comparison (S <= 4): an integer division per trip costs more than the
          // sums; out-of-range quads carry zeros
          const int si = (qq >= EQ) + (qq >= 2 * EQ) + (qq >= 3 * EQ);
#pragma unroll
          for (int k = 0; k < S; ++k)
            if (si == k) { sg[k] += gq; sn[k] += nv[u]; }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < S; ++k) {
      const float tg = GRAD ? warp_sum(sg[k]) : 0.f, tn = NM ? warp_sum(sn[k]) : 0.f;
      if (k < ns) {
        a0 += (double)sqrtf(tg);
        a1 += (double)sqrtf(tn);
      }
    }
  }
  // one atomic pair per block
  __shared__ double red[2][8];
  if (lane == 0) {
    red[0][threadIdx.x >> 5] = a0;
    red[1][threadIdx.x >> 5] = a1;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double v = 0.0;
    for (int w = 0; w < warps_per_block; ++w) v += red[threadIdx.x][w];
    if (v != 0.0) atomicAdd(acc + threadIdx.x, v);
  }
}

// acc = {sum ||grad_b||, sum ||noise_b||, ticket}: the last block to finish zeroes it for the next corrector step, so
// a captured CUDA graph needs no separate memset launch.
template <int U>
__global__ void __launch_bounds__(256)
corrector_update_kernel(const float4* __restrict__ x, const float4* __restrict__ grad, const float* __restrict__ t,
                        const float4* __restrict__ noise, double* __restrict__ acc,
                        const float* __restrict__ alphas, float4* __restrict__ x_out, float4* __restrict__ x_mean_out,
                        uint32_t n_quads, FastDiv eqd, SdeP s, float T, float target_snr, double inv_global_batch,
                        uint64_t seed, uint64_t draw, const uint64_t* draw_dev, uint64_t quad_offset, Impute im,
                        int reset_acc) {
  if (draw_dev) draw += *draw_dev;
  const float grad_norm = (float)(acc[0] * inv_global_batch);
  const float noise_norm = (float)(acc[1] * inv_global_batch);
  const float r = target_snr * noise_norm / grad_norm;
  const float base = r * r * 2.f;
  const float2 icoef = impute_coef(im, s);
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t q0 = blockIdx.x * blockDim.x + threadIdx.x; q0 < n_quads; q0 += U * stride) {
    uint32_t qs[U];
    float4 xv[U], gv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      qs[u] = q0 + u * stride;
      if (qs[u] < n_quads) {
        xv[u] = __ldcs(x + qs[u]);
        gv[u] = __ldcs(grad + qs[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t q = qs[u];
      if (q >= n_quads) break;
      const uint32_t b = fdiv(q, eqd);
      float alpha = 1.f;
      if (alphas != nullptr) {
        // (t * (N - 1) / T).long(): fp32 product, fp32 divide, truncate (sde_helper2.py:57)
        const int idx = (int)(__ldg(t + b) * (float)(s.N - 1) / T);
        alpha = __ldg(alphas + min(max(idx, 0), s.N - 1));
      }
      const float step = base * alpha;
      const float ns = sqrtf(step * 2.f);
      const float4 z = noise ? noise[q] : philox_normal4(seed, draw, quad_offset + (uint64_t)q);
      const float4 mean = make_float4(xv[u].x + step * gv[u].x, xv[u].y + step * gv[u].y, xv[u].z + step * gv[u].z,
                                      xv[u].w + step * gv[u].w);
      const float4 out = make_float4(mean.x + ns * z.x, mean.y + ns * z.y, mean.z + ns * z.z, mean.w + ns * z.w);
      if (x_mean_out) __stcs(x_mean_out + q, mean);
      __stcs(x_out + q, apply_impute(im, icoef, out, q, q - b * eqd.d));
    }
  }
  if (reset_acc) {
    __syncthreads();  // every thread of this block has read acc
    if (threadIdx.x == 0) {
      unsigned int* ticket = reinterpret_cast<unsigned int*>(acc + 2);
      __threadfence();
      if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
        acc[0] = 0.0;
        acc[1] = 0.0;
        *ticket = 0u;
      }
    }
  }
}

// standalone imputation (first step / finishing): x_out = imputed(x) ; final=1 writes the CLEAN latent
__global__ void __launch_bounds__(256)
impute_kernel(const float4* __restrict__ x, float4* __restrict__ x_out, uint32_t n_quads, FastDiv eqd, SdeP s,
              Impute im) {
  const float2 icoef = impute_coef(im, s);
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += gridDim.x * blockDim.x) {
    const uint32_t b = fdiv(q, eqd);
    x_out[q] = apply_impute(im, icoef, x[q], q, q - b * eqd.d);
  }
}

__global__ void __launch_bounds__(256)
randn_kernel(float* __restrict__ out, int64_t n_quads, uint64_t seed, uint64_t draw, uint64_t quad_offset,
             float scale) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n_quads;
       q += (int64_t)gridDim.x * blockDim.x) {
    float4 z = philox_normal4(seed, draw, quad_offset + (uint64_t)q);
    z.x *= scale; z.y *= scale; z.z *= scale; z.w *= scale;
    *reinterpret_cast<float4*>(out + q * 4) = z;
  }
}

// ------------------------------------------------------------------------------ DSM
// t_b = u_b*(T-eps)+eps ; x~ = mean_coef(t_b)*x0 + std(t_b)*z.  Writes x~, z (kept for the loss), t, std, g2.
__global__ void __launch_bounds__(256)
dsm_perturb_kernel(const float* __restrict__ x0, const float* __restrict__ u_in, const float* __restrict__ z_in,
                   float* __restrict__ xt, float* __restrict__ z_out, float* __restrict__ t_out,
                   float* __restrict__ std_out, float* __restrict__ g2_out, int64_t n_quads, int E, SdeP s, float T,
                   float eps, uint64_t seed, uint64_t draw_u, uint64_t draw_z, const uint64_t* draw_dev,
                   uint64_t sample_offset) {
  if (draw_dev) {  // CUDA-graph replay: the per-step draw id lives in device memory (sbm_train_tick advances it)
    draw_u += *draw_dev;
    draw_z += *draw_dev;
  }
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n_quads;
       q += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i4 = q * 4;
    const int b = (int)(i4 / E);
    const float u = u_in ? __ldg(u_in + b) : philox_uniform(seed, draw_u, sample_offset + (uint64_t)b);
    const float t = u * (T - eps) + eps;
    float mc, sd;
    sde_marginal(s, t, mc, sd);
    const float4 xv = *reinterpret_cast<const float4*>(x0 + i4);
    const float4 z = z_in ? *reinterpret_cast<const float4*>(z_in + i4)
                          : philox_normal4(seed, draw_z, sample_offset * (uint64_t)(E >> 2) + (uint64_t)q);
    *reinterpret_cast<float4*>(xt + i4) =
        make_float4(mc * xv.x + sd * z.x, mc * xv.y + sd * z.y, mc * xv.z + sd * z.z, mc * xv.w + sd * z.w);
    if (z_out) *reinterpret_cast<float4*>(z_out + i4) = z;
    if (i4 == (int64_t)b * E) {
      t_out[b] = t;
      std_out[b] = sd;
      if (g2_out) {
        float dc, g;
        sde_drift_diff(s, t, dc, g);
        g2_out[b] = g * g;
      }
    }
  }
}

// loss_b = red_CHW(term^2) [* g2_b] ; loss = mean_b.  term = score*std+z  (mode 0)  or  score + z/std (mode 1).
// Also writes dloss/dscore (unit upstream gradient).  One warp per sample; fp64 accumulation of the batch mean.
__global__ void __launch_bounds__(256)
dsm_loss_kernel(const float* __restrict__ score, const float* __restrict__ z, const float* __restrict__ std,
                const float* __restrict__ g2, float* __restrict__ dscore, double* __restrict__ loss_acc, int B,
                int E, int mode, int reduce_mean, double inv_global_batch) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int EQ = E >> 2;
  const float red_scale = reduce_mean ? 1.f / (float)E : 0.5f;
  double acc = 0.0;
  for (int b = blockIdx.x * warps_per_block + (threadIdx.x >> 5); b < B; b += gridDim.x * warps_per_block) {
    const float sd = __ldg(std + b);
    const float w = mode == 1 ? __ldg(g2 + b) : 1.f;
    // d loss / d score = (1/B) * w * red_scale * 2 * term * dterm/dscore
    const float gscale = (float)inv_global_batch * w * red_scale * 2.f * (mode == 0 ? sd : 1.f);
    float sum = 0.f;
    for (int q = lane; q < EQ; q += 32) {
      const int64_t i4 = ((int64_t)b * EQ + q) * 4;
      const float4 sv = *reinterpret_cast<const float4*>(score + i4);
      const float4 zv = *reinterpret_cast<const float4*>(z + i4);
      float4 term;
      if (mode == 0) term = make_float4(sv.x * sd + zv.x, sv.y * sd + zv.y, sv.z * sd + zv.z, sv.w * sd + zv.w);
      else term = make_float4(sv.x + zv.x / sd, sv.y + zv.y / sd, sv.z + zv.z / sd, sv.w + zv.w / sd);
      sum += term.x * term.x + term.y * term.y + term.z * term.z + term.w * term.w;
      if (dscore)
        *reinterpret_cast<float4*>(dscore + i4) =
            make_float4(gscale * term.x, gscale * term.y, gscale * term.z, gscale * term.w);
    }
    sum = warp_sum(sum);
    acc += (double)(sum * red_scale * w);
  }
  if (lane == 0 && acc != 0.0) atomicAdd(loss_acc, acc * inv_global_batch);
}

__global__ void scale_kernel(const float* __restrict__ in, const float* __restrict__ scalar, float* __restrict__ out,
                             int64_t n) {
  const float s = *scalar;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[i] * s;
}
__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)in[i];
}

__global__ void sampler_tick_kernel(const float* __restrict__ ts, int n_ts, int* step, unsigned long long* draw,
                                    float* __restrict__ t_vec, int B, float* t_next, int advance,
                                    unsigned long long draws_per_step) {
  const int st = min(*step, n_ts - 1);
  const float t = ts[st];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += gridDim.x * blockDim.x) t_vec[i] = t;
  // every block has read *step before the single writer below may bump it: enforce with a grid of ONE block
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (t_next) *t_next = ts[min(st + 1, n_ts - 1)];
    if (advance) {
      *step = st + 1;
      if (draw) *draw += draws_per_step;
    }
  }
}

// ------------------------------------------------------------------------------ legacy annealed-Langevin evaluators
// eval_lat_celeba_hq_all.py:268-275 and fid_upd10.py:279-290: for every modality channel m that is NOT observed
//   x' = x + a[m] * score + b[m] * noise          (observed channels are copied through)
// with a[m] = er[m] sigma_i^2 / sigma_last^2 / sigma_i, b[m] = c[m] sqrt(2 er[m] sigma_i^2 / sigma_last^2) for the
// annealed sampler and a = lr1 (i+1)/n_comp, b = lr2 for the fixed-step one.  12 B / element like the predictor.
struct ModCoef {
  float a[32], b[32];
};
__global__ void __launch_bounds__(256)
langevin_axpy_kernel(const float4* __restrict__ x, const float4* __restrict__ score, const float4* __restrict__ noise,
                     float4* __restrict__ x_out, uint32_t n_quads, FastDiv eqd, FastDiv ddq, uint32_t obs_mask,
                     ModCoef cf, uint64_t seed, uint64_t draw, const uint64_t* draw_dev, uint64_t quad_offset) {
  if (draw_dev) draw += *draw_dev;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < n_quads; q += gridDim.x * blockDim.x) {
    const uint32_t b = fdiv(q, eqd);
    const uint32_t m = fdiv(q - b * eqd.d, ddq);
    float4 xv = __ldcs(x + q);
    if (!((obs_mask >> m) & 1u)) {
      const float4 sv = __ldcs(score + q);
      const float4 z = noise ? noise[q] : philox_normal4(seed, draw, quad_offset + (uint64_t)q);
      const float a = cf.a[m], bb = cf.b[m];
      xv.x = xv.x + a * sv.x + bb * z.x;
      xv.y = xv.y + a * sv.y + bb * z.y;
      xv.z = xv.z + a * sv.z + bb * z.z;
      xv.w = xv.w + a * sv.w + bb * z.w;
    }
    __stcs(x_out + q, xv);
  }
}

// ------------------------------------------------------------------------------ classifier / EBM guidance glue
// new_x = cat(x[:, m1], x[:, m2]).view(B, 2*DD) (sde_helper2.py:70-71, 288-289) as the bf16 GEMM operand of the energy
// net: out[b][0:DD] = x[b][m1], out[b][DD:2DD] = x[b][m2], row stride ld (padding zeroed).
__global__ void __launch_bounds__(256)
guidance_gather_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int B, int M, int DD, int m1,
                       int m2, int ld) {
  const int64_t total = (int64_t)B * ld;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % ld);
    const int64_t b = idx / ld;
    float v = 0.f;
    if (j < 2 * DD) v = x[(b * M + (j < DD ? m1 : m2)) * DD + (j < DD ? j : j - DD)];
    out[idx] = __float2bfloat16_rn(v);
  }
}
// score[:, m1] -= cl_s * g[:, 0:DD]; score[:, m2] -= cl_s * g[:, DD:2DD]  (sde_helper2.py:75-76, 293-294); m < 0 skips
// that half (train_poly_unet_cont.py:87 updates the predicted modality only).  g: fp32 rows of stride ldg.
__global__ void __launch_bounds__(256)
guidance_apply_kernel(float* __restrict__ score, const float* __restrict__ g, int B, int M, int DD, int m1, int m2,
                      int64_t ldg, float cl_s) {
  const int64_t total = (int64_t)B * 2 * DD;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % (2 * DD));
    const int64_t b = idx / (2 * DD);
    const int m = j < DD ? m1 : m2;
    if (m < 0) continue;
    score[(b * M + m) * DD + (j < DD ? j : j - DD)] -= cl_s * g[b * ldg + j];
  }
}

// quads in flight per thread of the predictor / update kernels and Philox chains per lane of the noise-norm kernel
// (A/B: SBM_SAMPLER_UNROLL = 2 | 4, SBM_NOISE_UNROLL = 4 | 8; defaults = the measured best, profiles/README.md)
static int g_unroll = [] { const char* e = getenv("SBM_SAMPLER_UNROLL"); return e ? atoi(e) : 2; }();
static int g_noise_unroll = [] { const char* e = getenv("SBM_NOISE_UNROLL"); return e ? atoi(e) : 4; }();

static int ew_grid(int64_t n_items) {
  const int64_t want = (n_items + 255) / 256;
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count() * 8));
}
// Grid of a grid-stride kernel = exactly one resident wave (SMs x blocks that fit per SM): with a fixed cap like 8
// blocks per SM a kernel whose registers allow only 5 runs 1.6 waves and idles ~20 % of the machine in the second one.
template <typename K>
static int wave_grid(K kernel, int64_t n_items, int per_thread = 1) {
  // occupancy per kernel FUNCTION (template instantiations of one kernel share the pointer type K, not the pointer)
  static std::atomic<const void*> keys[16];
  static std::atomic<int> vals[16];
  int per_sm = 0;
  for (int i = 0; i < 16; ++i) {
    const void* k = keys[i].load(std::memory_order_acquire);
    if (k == (const void*)kernel) { per_sm = vals[i].load(std::memory_order_relaxed); break; }
    if (k == nullptr) {
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm <= 0) per_sm = 4;
      const void* expect = nullptr;
      vals[i].store(per_sm, std::memory_order_relaxed);  // a racing thread stores the same value for the same kernel
      if (!keys[i].compare_exchange_strong(expect, (const void*)kernel, std::memory_order_release) &&
          expect != (const void*)kernel)
        continue;  // slot taken by another kernel meanwhile: keep looking (per_sm is already known)
      break;
    }
  }
  if (per_sm <= 0) per_sm = 4;
  const int64_t want = (n_items + 256 * per_thread - 1) / (256 * per_thread);
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sm_count() * per_sm));
}
static int check_latent(const sbm_latent_shape* ls, const char* who) {
  SBM_CHECK_ARG(ls && ls->batch > 0 && ls->mods > 0 && ls->mods <= 32 && ls->dd > 0, "%s: bad latent shape", who);
  SBM_CHECK_ARG(ls->dd % 4 == 0, "%s: D*D = %d must be a multiple of 4 (vectorised latent access)", who, ls->dd);
  SBM_CHECK_ARG((int64_t)ls->batch * ls->mods * ls->dd / 4 < (int64_t(1) << 31),
                "%s: more than 2^31 latent quads in one call (shard the batch)", who);
  return 0;
}
static SdeP to_sdep(const sbm_sde* s) { return SdeP{s->kind, s->b0, s->b1, s->N}; }
static Impute to_impute(const sbm_impute* im, int dd) {
  Impute r;
  r.z_obs = im ? im->z_obs : nullptr;
  r.mask = im ? im->obs_mask : 0u;
  r.noise_obs = im ? im->noise_obs : 0;
  r.t_next = im ? im->t_next : 0.f;
  r.t_next_dev = im ? im->t_next_dev : nullptr;
  r.ddq = make_fastdiv((uint32_t)(dd / 4));
  if (r.mask == 0u) r.z_obs = nullptr;
  return r;
}

template <int S, bool GRAD, int NM, int NU>
static int launch_norms_u(const sbm_latent_shape* ls, const float* grad, const float* noise, const sbm_rng* rng,
                          double* acc2, void* stream) {
  const int EQ = ls->mods * ls->dd / 4;
  const int groups = (ls->batch + S - 1) / S;
  // a group moves S*EQ quads; size the grid by warp trips so that a small batch still spreads over the SMs
  const int blocks = wave_grid(corrector_norms_kernel<S, GRAD, NM, NU>, (int64_t)groups * 32);
  corrector_norms_kernel<S, GRAD, NM, NU><<<blocks, 256, 0, (cudaStream_t)stream>>>(
      (const float4*)grad, (const float4*)noise, acc2, ls->batch, EQ, rng ? rng->seed : 0, rng ? rng->draw : 0,
      rng ? rng->draw_dev : nullptr, rng ? rng->sample_offset * (uint64_t)EQ : 0);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_s();
  return 0;
}
template <int S, bool GRAD, int NM>
static int launch_norms(const sbm_latent_shape* ls, const float* grad, const float* noise, const sbm_rng* rng,
                        double* acc2, void* stream) {
  if (!GRAD && g_noise_unroll == 8) return launch_norms_u<S, GRAD, NM, (GRAD ? 4 : 8)>(ls, grad, noise, rng, acc2, stream);
  return launch_norms_u<S, GRAD, NM, 4>(ls, grad, noise, rng, acc2, stream);
}
template <bool GRAD, int NM>
static int dispatch_norms(const sbm_latent_shape* ls, const float* grad, const float* noise, const sbm_rng* rng,
                          double* acc2, void* stream) {
  const int EQ = ls->mods * ls->dd / 4;
  const int S = (EQ % 32 == 0) ? 1 : ((2 * EQ) % 32 == 0 ? 2 : ((4 * EQ) % 32 == 0 ? 4 : 1));
  if (S == 1) return launch_norms<1, GRAD, NM>(ls, grad, noise, rng, acc2, stream);
  if (S == 2) return launch_norms<2, GRAD, NM>(ls, grad, noise, rng, acc2, stream);
  return launch_norms<4, GRAD, NM>(ls, grad, noise, rng, acc2, stream);
}

static int launch_predictor(const sbm_latent_shape* ls, const sbm_sde* sde, const float* x, const float* score,
                            const float* t, const float* noise, float* x_out, float* x_mean_out,
                            int32_t probability_flow, const sbm_rng* rng, const sbm_impute* impute, int rd,
                            const float* table, void* stream) {
  const int E = ls->mods * ls->dd;
  const int64_t nq = (int64_t)ls->batch * E / 4;
#define SBM_LAUNCH_PREDICTOR(U)                                                                                       \
  predictor_kernel<U><<<wave_grid(predictor_kernel<U>, nq, U), 256, 0, (cudaStream_t)stream>>>(                        \
      (const float4*)x, (const float4*)score, t, (const float4*)noise, (float4*)x_out, (float4*)x_mean_out,           \
      (uint32_t)nq, make_fastdiv((uint32_t)(E / 4)), to_sdep(sde), probability_flow, rng ? rng->seed : 0,             \
      rng ? rng->draw : 0, rng ? rng->draw_dev : nullptr, rng ? rng->sample_offset * (uint64_t)(E / 4) : 0,           \
      to_impute(impute, ls->dd), rd, table, sde->T)
  if (g_unroll == 4) SBM_LAUNCH_PREDICTOR(4); else SBM_LAUNCH_PREDICTOR(2);
#undef SBM_LAUNCH_PREDICTOR
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_s();
  return 0;
}

}  // namespace sbm

using namespace sbm;

extern "C" {

int sbm_randn(float* out, int64_t n, uint64_t seed, uint64_t draw, uint64_t elem_offset, float scale, void* stream) {
  SBM_CHECK_ARG(out && n > 0 && n % 4 == 0 && elem_offset % 4 == 0, "sbm_randn: n and offset must be multiples of 4");
  randn_kernel<<<ew_grid(n / 4), 256, 0, (cudaStream_t)stream>>>(out, n / 4, seed, draw, elem_offset / 4, scale);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_s();
  return 0;
}

int sbm_dropout(void* x, int32_t dtype, int64_t ld, int64_t rows, int32_t C, float p, const sbm_rng* rng,
                void* stream) {
  SBM_CHECK_ARG(x && rows > 0 && C > 0 && rng, "sbm_dropout: bad args");
  SBM_CHECK_ARG(p >= 0.f && p < 1.f, "sbm_dropout: p = %g must be in [0, 1)", (double)p);
  SBM_CHECK_ARG(dtype == SBM_F32 || dtype == SBM_BF16, "sbm_dropout: dtype");
  const int c8 = (C + 7) / 8;
  SBM_CHECK_ARG(ld >= (int64_t)c8 * 8 && ld % 8 == 0, "sbm_dropout: row stride %lld must cover C rounded up to 8",
                (long long)ld);
  const int64_t n_oct = rows * c8;
  SBM_CHECK_ARG(n_oct < (int64_t(1) << 32), "sbm_dropout: more than 2^32 channel octets");
  const float inv_keep = 1.f / (1.f - p);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SBM_BF16)
    dropout_kernel<__nv_bfloat16><<<wave_grid(dropout_kernel<__nv_bfloat16>, n_oct), 256, 0, st>>>(
        (__nv_bfloat16*)x, ld, n_oct, make_fastdiv((uint32_t)c8), C, p, inv_keep, rng->seed, rng->draw, rng->draw_dev);
  else
    dropout_kernel<float><<<wave_grid(dropout_kernel<float>, n_oct), 256, 0, st>>>(
        (float*)x, ld, n_oct, make_fastdiv((uint32_t)c8), C, p, inv_keep, rng->seed, rng->draw, rng->draw_dev);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_s();
  return 0;
}

int sbm_sampler_tick(const float* ts, int32_t n_ts, int32_t* step, uint64_t* draw, float* t_vec, int32_t B,
                     float* t_next, int32_t advance, uint64_t draws_per_step, void* stream) {
  SBM_CHECK_ARG(ts && step && t_vec && n_ts > 0 && B > 0, "sbm_sampler_tick: bad args");
  sampler_tick_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ts, n_ts, step, (unsigned long long*)draw, t_vec, B, t_next,
                                                            advance, draws_per_step);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_s();
  return 0;
}

__global__ void train_tick_kernel(int* step, unsigned long long* draw, unsigned long long draw_inc) {
  if (step) *step += 1;
  if (draw) *draw += draw_inc;
}
int sbm_train_tick(int32_t* step_dev, uint64_t* draw_dev, uint64_t draw_inc, void* stream) {
  SBM_CHECK_ARG(step_dev || draw_dev, "sbm_train_tick: nothing to advance");
  train_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev, (unsigned long long*)draw_dev, draw_inc);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_s();
  return 0;
}

int sbm_predictor_step(const sbm_latent_shape* ls, const sbm_sde* sde, const float* x, const float* score,
                       const float* t, const float* noise, float* x_out, float* x_mean_out, int32_t probability_flow,
                       const sbm_rng* rng, const sbm_impute* impute, void* stream) {
  if (check_latent(ls, "sbm_predictor_step")) return 1;
  SBM_CHECK_ARG(sde && x && score && t && x_out, "sbm_predictor_step: null pointer");
  SBM_CHECK_ARG(noise || rng || probability_flow, "sbm_predictor_step: need injected noise or an rng");
  return launch_predictor(ls, sde, x, score, t, noise, x_out, x_mean_out, probability_flow, rng, impute, 0, nullptr,
                          stream);
}

int sbm_rd_predictor_step(const sbm_latent_shape* ls, const sbm_sde* sde, const float* x, const float* score,
                          const float* t, const float* table, const float* noise, float* x_out, float* x_mean_out,
                          int32_t probability_flow, const sbm_rng* rng, const sbm_impute* impute, void* stream) {
  if (check_latent(ls, "sbm_rd_predictor_step")) return 1;
  SBM_CHECK_ARG(sde && x && score && t && x_out, "sbm_rd_predictor_step: null pointer");
  SBM_CHECK_ARG(noise || rng || probability_flow, "sbm_rd_predictor_step: need injected noise or an rng");
  SBM_CHECK_ARG(table || sde->kind == SBM_SDE_SUBVP,
                "sbm_rd_predictor_step: VPSDE needs the discrete_betas table, VESDE the discrete_sigmas table");
  return launch_predictor(ls, sde, x, score, t, noise, x_out, x_mean_out, probability_flow, rng, impute, 1, table,
                          stream);
}

int sbm_corrector_norms(const sbm_latent_shape* ls, const float* grad, const float* noise, const sbm_rng* rng,
                        double* acc2, void* stream) {
  if (check_latent(ls, "sbm_corrector_norms")) return 1;
  SBM_CHECK_ARG(grad && acc2, "sbm_corrector_norms: null pointer");
  if (noise) return dispatch_norms<true, 1>(ls, grad, noise, nullptr, acc2, stream);
  if (rng) return dispatch_norms<true, 2>(ls, grad, nullptr, rng, acc2, stream);
  return dispatch_norms<true, 0>(ls, grad, nullptr, nullptr, acc2, stream);  // acc2[1] comes from sbm_noise_norm
}

int sbm_noise_norm(const sbm_latent_shape* ls, const sbm_rng* rng, double* acc2, void* stream) {
  if (check_latent(ls, "sbm_noise_norm")) return 1;
  SBM_CHECK_ARG(rng && acc2, "sbm_noise_norm: null pointer");
  return dispatch_norms<false, 2>(ls, nullptr, nullptr, rng, acc2, stream);
}

int sbm_corrector_update(const sbm_latent_shape* ls, const sbm_sde* sde, const float* x, const float* grad,
                         const float* t, const float* noise, double* acc2, const float* alphas, float* x_out,
                         float* x_mean_out, float target_snr, int64_t global_batch, const sbm_rng* rng,
                         const sbm_impute* impute, int32_t reset_acc, void* stream) {
  if (check_latent(ls, "sbm_corrector_update")) return 1;
  SBM_CHECK_ARG(sde && x && grad && t && acc2 && x_out && (noise || rng), "sbm_corrector_update: null pointer");
  SBM_CHECK_ARG(global_batch >= ls->batch, "sbm_corrector_update: global_batch < local batch");
  const int E = ls->mods * ls->dd;
  const int64_t nq = (int64_t)ls->batch * E / 4;
#define SBM_LAUNCH_UPDATE(U)                                                                                          \
  corrector_update_kernel<U><<<wave_grid(corrector_update_kernel<U>, nq, U), 256, 0, (cudaStream_t)stream>>>(          \
      (const float4*)x, (const float4*)grad, t, (const float4*)noise, acc2, alphas, (float4*)x_out,                   \
      (float4*)x_mean_out, (uint32_t)nq, make_fastdiv((uint32_t)(E / 4)), to_sdep(sde), sde->T, target_snr,           \
      1.0 / (double)global_batch, rng ? rng->seed : 0, rng ? rng->draw : 0, rng ? rng->draw_dev : nullptr,            \
      rng ? rng->sample_offset * (uint64_t)(E / 4) : 0, to_impute(impute, ls->dd), reset_acc)
  if (g_unroll == 4) SBM_LAUNCH_UPDATE(4); else SBM_LAUNCH_UPDATE(2);
#undef SBM_LAUNCH_UPDATE
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_s();
  return 0;
}

int sbm_langevin_axpy_step(const sbm_latent_shape* ls, const float* x, const float* score, const float* noise,
                           const float* coef_a, const float* coef_b, uint32_t obs_mask, float* x_out,
                           const sbm_rng* rng, void* stream) {
  if (check_latent(ls, "sbm_langevin_axpy_step")) return 1;
  SBM_CHECK_ARG(x && score && x_out && coef_a && coef_b && (noise || rng), "sbm_langevin_axpy_step: null pointer");
  ModCoef cf;
  for (int m = 0; m < 32; ++m) {  // HOST arrays of ls->mods coefficients
    cf.a[m] = m < ls->mods ? coef_a[m] : 0.f;
    cf.b[m] = m < ls->mods ? coef_b[m] : 0.f;
  }
  const int E = ls->mods * ls->dd;
  const int64_t nq = (int64_t)ls->batch * E / 4;
  langevin_axpy_kernel<<<wave_grid(langevin_axpy_kernel, nq), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)x, (const float4*)score, (const float4*)noise, (float4*)x_out, (uint32_t)nq,
      make_fastdiv((uint32_t)(E / 4)), make_fastdiv((uint32_t)(ls->dd / 4)), obs_mask, cf, rng ? rng->seed : 0,
      rng ? rng->draw : 0, rng ? rng->draw_dev : nullptr, rng ? rng->sample_offset * (uint64_t)(E / 4) : 0);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_s();
  return 0;
}

int sbm_guidance_gather(const sbm_latent_shape* ls, const float* x, int32_t m1, int32_t m2, void* out_bf16, int32_t ld,
                        void* stream) {
  if (check_latent(ls, "sbm_guidance_gather")) return 1;
  SBM_CHECK_ARG(x && out_bf16 && m1 >= 0 && m1 < ls->mods && m2 >= 0 && m2 < ls->mods && ld >= 2 * ls->dd,
                "sbm_guidance_gather: bad args");
  guidance_gather_kernel<<<ew_grid((int64_t)ls->batch * ld), 256, 0, (cudaStream_t)stream>>>(
      x, (__nv_bfloat16*)out_bf16, ls->batch, ls->mods, ls->dd, m1, m2, ld);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_s();
  return 0;
}

int sbm_guidance_apply(const sbm_latent_shape* ls, float* score, const float* grad, int64_t ldg, int32_t m1, int32_t m2,
                       float cl_s, void* stream) {
  if (check_latent(ls, "sbm_guidance_apply")) return 1;
  SBM_CHECK_ARG(score && grad && m1 < ls->mods && m2 < ls->mods && (m1 >= 0 || m2 >= 0) && ldg >= 2 * ls->dd,
                "sbm_guidance_apply: bad args");
  guidance_apply_kernel<<<ew_grid((int64_t)ls->batch * 2 * ls->dd), 256, 0, (cudaStream_t)stream>>>(
      score, grad, ls->batch, ls->mods, ls->dd, m1, m2, ldg, cl_s);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_s();
  return 0;
}

int sbm_impute_observed(const sbm_latent_shape* ls, const sbm_sde* sde, const float* x, float* x_out,
                        const sbm_impute* impute, void* stream) {
  if (check_latent(ls, "sbm_impute_observed")) return 1;
  SBM_CHECK_ARG(sde && x && x_out && impute && impute->z_obs, "sbm_impute_observed: null pointer");
  const int E = ls->mods * ls->dd;
  const int64_t nq = (int64_t)ls->batch * E / 4;
  impute_kernel<<<ew_grid(nq), 256, 0, (cudaStream_t)stream>>>((const float4*)x, (float4*)x_out, (uint32_t)nq,
                                                               make_fastdiv((uint32_t)(E / 4)), to_sdep(sde),
                                                               to_impute(impute, ls->dd));
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_s();
  return 0;
}

int sbm_dsm_perturb(const sbm_latent_shape* ls, const sbm_sde* sde, const float* x0, const float* u, const float* z,
                    float* xt, float* z_out, float* t_out, float* std_out, float* g2_out, float eps,
                    const sbm_rng* rng, void* stream) {
  if (check_latent(ls, "sbm_dsm_perturb")) return 1;
  SBM_CHECK_ARG(sde && x0 && xt && t_out && std_out, "sbm_dsm_perturb: null pointer");
  SBM_CHECK_ARG((u && z) || rng, "sbm_dsm_perturb: need injected (u, z) or an rng");
  const int E = ls->mods * ls->dd;
  const int64_t nq = (int64_t)ls->batch * E / 4;
  dsm_perturb_kernel<<<wave_grid(dsm_perturb_kernel, nq), 256, 0, (cudaStream_t)stream>>>(
      x0, u, z, xt, z_out, t_out, std_out, g2_out, nq, E, to_sdep(sde), sde->T, eps, rng ? rng->seed : 0,
      rng ? rng->draw : 0, rng ? rng->draw + 1 : 0, rng ? rng->draw_dev : nullptr, rng ? rng->sample_offset : 0);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_s();
  return 0;
}

int sbm_dsm_loss(const sbm_latent_shape* ls, const float* score, const float* z, const float* std, const float* g2,
                 float* dscore, double* loss_acc, int32_t likelihood_weighting, int32_t reduce_mean,
                 int64_t global_batch, void* stream) {
  if (check_latent(ls, "sbm_dsm_loss")) return 1;
  SBM_CHECK_ARG(score && z && std && loss_acc, "sbm_dsm_loss: null pointer");
  SBM_CHECK_ARG(!likelihood_weighting || g2, "sbm_dsm_loss: likelihood weighting needs g2");
  const int E = ls->mods * ls->dd;
  const int blocks = std::max(1, std::min((ls->batch + 7) / 8, sm_count() * 8));
  dsm_loss_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(score, z, std, g2, dscore, loss_acc, ls->batch, E,
                                                            likelihood_weighting ? 1 : 0, reduce_mean,
                                                            1.0 / (double)global_batch);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_s();
  return 0;
}

int sbm_scale_by_scalar(const float* in, const float* scalar_dev, float* out, int64_t n, void* stream) {
  SBM_CHECK_ARG(in && scalar_dev && out && n > 0, "sbm_scale_by_scalar: bad args");
  scale_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(in, scalar_dev, out, n);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_s();
  return 0;
}

int sbm_f64_to_f32(const double* in, float* out, int32_t n, void* stream) {
  SBM_CHECK_ARG(in && out && n > 0, "sbm_f64_to_f32: bad args");
  f64_to_f32_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(in, out, n);
  SBM_CUDA_OK(cudaGetLastError());
  count_launch_s();
  return 0;
}

}  // extern "C"
