/*
 * libsbmae_b200 — C ABI of the B200-native SBM-AE latent score-model hot path.
 *
 * The reference (DanielMitiku/score_based_multimodal_autoencoder) has no FFI: its
 * boundary is Python duck typing between the experiment scripts and three library
 * modules (sde_helper2.py, unet_model.py, unet_openai.py).  Every entry point below
 * names the reference statement(s) it replaces (file:line under /root/reference).
 * The Python host (score_based_multimodal_autoencoder_b200/*.py) mirrors the
 * reference's classes/functions and calls these through ctypes.
 *
 * Conventions
 *  - all pointers are DEVICE pointers unless stated otherwise; no allocation, no
 *    ownership transfer, no hidden global state except immutable caches -- and the
 *    process-wide MEASUREMENT switches (sbm_conv_force_single_cta, sbm_conv_force_direct_epilogue,
 *    sbm_conv_epilogue_static, sbm_conv_pixel_major, the SBM_* environment knobs read once at load), which select between
 *    kernels that compute the same result (bit-identical or equal up to fp32 summation order)
 *    and exist for A/B timing only: a production caller never touches them;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*) and
 *    capturable into a CUDA graph;
 *  - return 0 on success; non-zero on error, message via sbm_last_error();
 *  - activations inside the score net are channels-last: pixel-major rows of
 *    `ld` elements (ld >= channels, multiple of 8), spatial extents powers of two
 *    (the reference pads to powers of two, unet_model.py:276-284).
 */
#ifndef SBMAE_B200_H_
#define SBMAE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* sbm_last_error(void);
int sbm_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
unsigned long long sbm_launch_count(void);

/* ------------------------------------------------------------------ convolutions
 * Implicit-GEMM convolution on the tcgen05 tensor cores (bf16 x bf16 -> fp32 in TMEM),
 * operands staged by TMA; replaces nn.Conv2d / nn.ConvTranspose2d / nn.Linear calls of
 * unet_model.py:30,33,103-110,132-133,157-160,208,222-227,272 and
 * unet_openai.py:185,207,253-268,322-324,421-425.
 */
enum { SBM_CONV_S1 = 0,     /* KHxKW, stride 1, padding (K-1)/2 ("same")            */
       SBM_CONV_S2 = 1,     /* KHxKW (4x4 or 3x3), stride 2, padding 1              */
       SBM_CONVT_4X4_S2 = 2 /* ConvTranspose2d 4x4 (or 3x3 + output_padding 1), stride 2, padding 1 */ };
enum { SBM_ACT_NONE = 0, SBM_ACT_GELU = 1, SBM_ACT_SILU = 2 };
enum { SBM_F32 = 0, SBM_BF16 = 1 };

typedef struct sbm_conv_args {
  int32_t kind, kh, kw;
  int32_t batch, h, w;      /* INPUT spatial extent (powers of two, <= 128)          */
  int32_t cin, cout;
  const void* x;            /* bf16 [batch,h,w,ldx]                                   */
  int64_t ldx;
  const void* wpk;          /* bf16 [kh*kw][cout][cin_pad], from sbm_pack_weight_bf16 */
  int32_t cin_pad;
  int32_t act;              /* applied after bias, before residual                    */
  const float* bias;        /* [cout] or NULL                                         */
  const void* residual;     /* output geometry, pixel stride ldr, dtype res_dtype; or NULL */
  int64_t ldr;
  void* out;                /* [batch,oh,ow,ldo] (or NCHW fp32 when out_nchw)         */
  int64_t ldo;
  int32_t out_dtype;        /* SBM_F32 / SBM_BF16                                     */
  int32_t out_nchw;         /* 1: write fp32 [batch,cout,oh,ow] (final layer)         */
  int32_t res_dtype;        /* SBM_F32 / SBM_BF16                                     */
  int32_t out2_preact;      /* 1: out2 receives the PRE-activation value (bias added, before act/residual)   */
  double* stats;            /* [batch][2] += (sum, sum of squares) of the written values, or NULL */
  void* out2;               /* optional second copy of the output in bf16 (pixel stride ldo2), or NULL */
  int64_t ldo2;
  const float* rowbias;     /* optional per-sample bias [batch][ld_rowbias] (unet_openai.py:303 `h + emb_out`), or NULL */
  int64_t ld_rowbias;
  /* GroupNorm(1, cin) of the input folded into this convolution (unet_model.py:106-110: GN -> conv): x is the RAW
   * (un-normalised) tensor, wpk / gn_tab come from sbm_conv_fold_groupnorm, gn_stats = (sum, sumsq) per sample of x
   * (e.g. the `stats` output of the producing convolution), gn_count = H*W*cin.  bias must be NULL (it is in gn_tab).
   * The epilogue computes rstd*(acc - mean*Sg[cls]) + Tb[cls]; cls tells which 3x3 taps see real pixels. */
  const double* gn_stats;
  const float* gn_tab;      /* [2][16][cout] */
  float gn_count;
  float gn_eps;
  /* optional split-K workspace: ws_elems floats (16-byte aligned), rows of ld_ws >= cout floats (multiple of 4), need
   * not be initialised.  When given and sbm_conv_splitk_plan(a) > 1 (a stride-1 or stride-2 layer of a few 256 x 256
   * output tiles with a long K loop: the low-resolution levels at small batch), the K loop is cut across the SM pairs:
   * slice s stores its partial accumulators into slab s of the workspace, a second kernel sums the slabs in slice order
   * (deterministic) and applies the epilogue.  Size: sbm_conv_splitk_ws_elems(a).  NULL: never split.  One buffer can
   * serve every call issued in stream order. */
  float* splitk_ws;
  int64_t ld_ws;
  int64_t ws_elems;
} sbm_conv_args;

int sbm_conv_igemm(const sbm_conv_args* a, void* stream);
/* number of K slices sbm_conv_igemm would use for this call when given a workspace (1 = it would not split) */
int sbm_conv_splitk_plan(const sbm_conv_args* a);
/* floats the split-K workspace of this call needs with rows of a->ld_ws floats (0 = the call would not split) */
int64_t sbm_conv_splitk_ws_elems(const sbm_conv_args* a);
/* A/B switch: 0 = never split along K (sbm_conv_splitk_plan then returns 1); default 1, also SBM_SPLITK */
int sbm_conv_splitk(int32_t on);
/* A/B switch for measurements: 1 = always use the single-CTA kernel instead of the CTA-pair (cta_group::2) one */
int sbm_conv_force_single_cta(int32_t on);
/* A/B switch: 1 = per-thread global stores in the CTA-pair kernel instead of the TMA-staged epilogue */
int sbm_conv_force_direct_epilogue(int32_t on);
/* A/B switch: 0 = always run the run-time tested epilogue loop of the CTA-pair kernel instead of the statically
 * compiled loop of the call's flag set (same arithmetic in the same order: bit-identical results); default 1 */
int sbm_conv_epilogue_static(int32_t on);
/* which kernel the calling thread's last sbm_conv_igemm used: N tile | CTA-pair << 16 | staged epilogue << 17 |
 * pixel-major tiling << 18 | statically compiled epilogue mode << 19 | split-K << 20 | K slices << 24 */
int sbm_conv_last_variant(void);
/* pixel-major tiling of stride-1 'same' convolutions (a tile = 128 samples at ONE output pixel, so taps that only read
 * zero padding there are skipped): -1 = decide by work estimate (default), 0 = never, 1 = whenever the CTA-pair kernel
 * runs the layer.  Results are bit-identical either way (the skipped products are exact zeros).  Returns the old mode */
int sbm_conv_pixel_major(int32_t mode);

/* Weight gradient of sbm_conv_igemm: dwpk[tap][o][i] += sum_pixels dy[p][o] * x[p shifted by tap][i]  (fp32; split-K
 * partial tiles are ADDED into dwpk by TMA reduce-add boxes, so the caller zeroes dwpk).  x = the forward input
 * operand (bf16), dy = gradient of the forward output (bf16, output geometry).  Backward of the nn.Conv2d / nn.ConvTranspose2d / nn.Linear weights under loss.backward()
 * (train_lat_celebhq_unet_cont2.py:98-100). */
typedef struct sbm_wgrad_args {
  int32_t kind, kh, kw;
  int32_t batch, h, w;      /* INPUT spatial extent of the forward convolution */
  int32_t cin, cout;
  const void* x;  int64_t ldx;
  const void* dy; int64_t lddy;
  float* dwpk;    int32_t cin_pad; int32_t reserved;
} sbm_wgrad_args;
int sbm_conv_wgrad(const sbm_wgrad_args* a, void* stream);
/* packed fp32 gradient [taps][rows][cols_pad] -> parameter layout: dst[tap*s_tap + row*s_row + col*s_col] */
int sbm_unpack_wgrad(const float* src, float* dst, int32_t taps, int32_t rows, int32_t cols, int32_t cols_pad,
                     int64_t s_tap, int64_t s_row, int64_t s_col, void* stream);

/* Multi-tensor form of sbm_unpack_wgrad: every listed packed gradient -> its parameter layout in ONE launch.  Same
 * descriptor as sbm_pack_weights_multi with the roles swapped: src = packed fp32 [taps][rows][cols_pad], dst = fp32
 * parameter-gradient layout addressed by (s_tap, s_row, s_col); tensor k owns blocks [first_block, first_block +
 * ceil(rows/32)*tiles_c), tiles_c = ceil(cols/32); sorted by first_block. */
struct sbm_pack_desc;
int sbm_unpack_wgrad_multi(const struct sbm_pack_desc* descs_dev, int32_t n_descs, int32_t n_blocks, int32_t max_taps,
                           void* stream);

/* GroupNorm(1,C) -> conv folding: dst = bf16 [kh*kw][rows][cols_pad] of w*gamma[col]; tab[0][cls][row] = sum over the
 * taps valid in border class cls of sum_col bf16(w*gamma), tab[1][cls][row] = same sum of w*beta, + bias[row].
 * cls bit0: tap row 0 valid (pixel row >= 1), bit1: tap row 2 valid, bit2 / bit3: same for columns. */
int sbm_conv_fold_groupnorm(const float* w, void* dst, float* tab, int32_t kh, int32_t kw, int32_t rows, int32_t cols,
                            int32_t cols_pad, int64_t s_tap, int64_t s_row, int64_t s_col, const float* gamma,
                            const float* beta, const float* bias, void* stream);
/* Multi-tensor form of sbm_pack_weight_bf16: re-packs every listed weight in ONE launch (a training step re-packs all
 * ~180 GEMM operands after the optimizer step).  descs_dev: DEVICE array sorted by first_block; tensor k owns blocks
 * [first_block, first_block + ceil(rows/32)*tiles_c), tiles_c = ceil(cols_pad/32);
 * taps_magic = taps > 1 ? ceil(2^32 / taps) : 0. */
typedef struct sbm_pack_desc {
  const float* src; void* dst;
  int32_t taps, rows, cols, cols_pad;
  int64_t s_tap, s_row, s_col;
  int32_t tiles_c, first_block;
  uint32_t taps_magic; int32_t reserved;
} sbm_pack_desc;
int sbm_pack_weights_multi(const sbm_pack_desc* descs_dev, int32_t n_descs, int32_t n_blocks, int32_t max_taps,
                           void* stream);
/* fp32 weights -> bf16 [taps][rows][cols_pad]; src element (tap,row,col) at
 * w[tap*s_tap + row*s_row + col*s_col]; optional per-column scale (GroupNorm gamma folding). */
int sbm_pack_weight_bf16(const float* w, void* dst, int32_t taps, int32_t rows, int32_t cols, int32_t cols_pad,
                         int64_t s_tap, int64_t s_row, int64_t s_col, void* stream);

/* ------------------------------------------------------------------ memory-bound score-net operators */
/* fp32 NCHW latent -> bf16 im2col rows [B*H*W, ldk], column (c*kh + i)*kw + j; feeds the stem conv
 * (unet_model.py:208,287; unet_openai.py:441) as a 1x1 sbm_conv_igemm. */
int sbm_stem_im2col(const float* x, void* a, int32_t B, int32_t C, int32_t H, int32_t W, int32_t kh, int32_t kw,
                    int32_t ldk, void* stream);
/* out = depthwise7x7(x) + bias[c] + cond[b][c]; stats[b] += (sum, sumsq).  unet_model.py:103,116-121.
 * x/out fp32 channels-last, w is the nn.Conv2d(groups=C) weight [C,1,7,7]. */
int sbm_dwconv7_fwd(const float* x, int64_t ldx, const float* w, const float* bias, const float* cond, int64_t ldc,
                    float* out, int64_t ldo, double* stats, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
/* same, output stored as bf16 (the statistics are those of the unrounded values): feeds a convolution that has the
 * following GroupNorm folded in (sbm_conv_fold_groupnorm).  Square power-of-two maps up to 16x16. */
int sbm_dwconv7_fwd_bf16(const float* x, int64_t ldx, const float* w, const float* bias, const float* cond,
                         int64_t ldc, void* out_bf16, int64_t ldo, double* stats, int32_t B, int32_t H, int32_t W,
                         int32_t C, void* stream);
/* same kernel, backward w.r.t. the input: out = depthwise7x7(dy, taps flipped) (+ addend, same layout as out) */
int sbm_dwconv7_bwd_input(const float* dy, int64_t lddy, const float* w, const float* addend, int64_t ldadd,
                          float* out, int64_t ldo, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
/* stats[b][g] += (sum, sumsq) of group g of sample b.  nn.GroupNorm statistics (unet_model.py:106,109,160,183;
 * unet_openai.py:10-12). */
int sbm_group_stats(const void* x, int32_t in_dtype, int64_t ldx, int32_t B, int32_t HW, int32_t C, int32_t G,
                    double* stats, void* stream);
/* y = act(GroupNorm(x; stats) * gamma + beta) (+ residual); writes `out` (out_dtype) and/or `out_f32`. */
int sbm_groupnorm_apply(const void* x, int32_t in_dtype, int64_t ldx, const double* stats, const float* gamma,
                        const float* beta, const float* residual, int64_t ldr, void* out, int32_t out_dtype,
                        int64_t ldo, float* out_f32, int64_t ldo_f32, int32_t B, int32_t HW, int32_t C, int32_t G,
                        float eps, int32_t act, void* stream);
/* the same with per-sample modulation: y = act(GroupNorm(x) * (1 + mod_scale[b][c]) + mod_shift[b][c]) +
 * post_add[b][c] (+ residual).  mod_scale / mod_shift (fp32 [B][ld_mod], both or neither): `use_scale_shift_norm` of
 * unet_openai.py:296-300; post_add (fp32 [B][ld_post]): the time embedding added AFTER the SiLU of Block 1 in
 * ResnetBlock (unet_model.py:82-87).  Null pointers: plain sbm_groupnorm_apply. */
int sbm_groupnorm_apply_mod(const void* x, int32_t in_dtype, int64_t ldx, const double* stats, const float* gamma,
                            const float* beta, const float* residual, int64_t ldr, void* out, int32_t out_dtype,
                            int64_t ldo, float* out_f32, int64_t ldo_f32, int32_t B, int32_t HW, int32_t C, int32_t G,
                            float eps, int32_t act, const float* mod_scale, const float* mod_shift, int64_t ld_mod,
                            const float* post_add, int64_t ld_post, void* stream);
/* nearest-neighbour 2x upsampling of a bf16 channels-last map (unet_openai.py:185) */
int sbm_upsample_nearest2x(const void* x, int64_t ldx, void* out, int64_t ldo, int32_t B, int32_t H, int32_t W,
                           int32_t C, void* stream);
/* sinusoidal embedding of t[B] -> bf16 [B, ld]; mode 0 = unet_model.py:40-47, mode 1 = unet_openai.py:66-83 */
int sbm_time_embed(const float* t, void* out_bf16, float* out_f32, int32_t B, int32_t dim, int32_t ld, int32_t mode,
                   void* stream);
/* LinearAttention core (unet_model.py:162-177) on qkv fp32 [B,n,ldq] (channels q|k|v, heads x 32) -> bf16 [B,n,ldo] */
int sbm_linear_attn_fwd(const float* qkv, int64_t ldq, void* out, int64_t ldo, int32_t B, int32_t n, int32_t heads,
                        float scale, void* stream);
/* the same on a bf16 qkv tensor (n = 64 or 256 positions): the to_qkv GEMM writes and the attention kernel reads half
 * the bytes; the soft-max arithmetic stays fp32 */
int sbm_linear_attn_fwd_bf16(const void* qkv, int64_t ldq, void* out, int64_t ldo, int32_t B, int32_t n, int32_t heads,
                             float scale, void* stream);
/* softmax attention core (unet_model.py:135-149; unet_openai.py:345-358): channel of (head,d) for q is
 * q_off + head*head_stride + d (k_off, v_off likewise); logits scaled by `scale`. */
int sbm_softmax_attn_fwd(const float* qkv, int64_t ldq, void* out, int64_t ldo, int32_t B, int32_t n, int32_t heads,
                         int32_t dh, int32_t q_off, int32_t k_off, int32_t v_off, int32_t head_stride, float scale,
                         void* stream);

/* ------------------------------------------------------------------ sampler / DSM (latent [B, mods, D, D] fp32) */
enum { SBM_SDE_VP = 0, SBM_SDE_SUBVP = 1, SBM_SDE_VE = 2 };
typedef struct sbm_latent_shape { int32_t batch, mods, dd; /* dd = D*D */ } sbm_latent_shape;
/* VPSDE/subVPSDE: (b0,b1) = (beta_min,beta_max); VESDE: (sigma_min,sigma_max).  sde_helper2.py:329-473 */
typedef struct sbm_sde { int32_t kind; float b0, b1; int32_t N; float T; } sbm_sde;
/* Philox4x32-10 stream: key = seed, counter = (global element index / 4, draw); sample_offset = index of this
 * shard's first sample inside the global batch (multi-GPU sampling draws the same numbers as one GPU would). */
typedef struct sbm_rng {
  uint64_t seed, draw, sample_offset;
  const uint64_t* draw_dev; /* optional DEVICE counter added to `draw` (lets one captured CUDA graph serve every step) */
} sbm_rng;
/* observed-modality imputation applied to the state a step WRITES (train_lat_celebhq_unet_cont2.py:293-303):
 * channel m with bit m of obs_mask set := noise_obs ? exp(lmc(t_next))*z + std(t_next)*z : z */
typedef struct sbm_impute {
  const float* z_obs; uint32_t obs_mask; int32_t noise_obs; float t_next;
  const float* t_next_dev; /* optional DEVICE scalar overriding t_next (CUDA-graph replay) */
} sbm_impute;
/* per-step device state of a graph-replayed sampler: t_vec[0..B) = ts[*step]; *t_next = ts[*step+1] (or ts[*step] when
 * last); then, when `advance` != 0: *step += 1 and *draw += draws_per_step.  One tiny kernel. */
int sbm_sampler_tick(const float* ts, int32_t n_ts, int32_t* step, uint64_t* draw, float* t_vec, int32_t B,
                     float* t_next, int32_t advance, uint64_t draws_per_step, void* stream);

int sbm_randn(float* out, int64_t n, uint64_t seed, uint64_t draw, uint64_t elem_offset, float scale, void* stream);
/* em_predictor, sde_helper2.py:45-52: x_mean = x + (f(x,t) - g^2 s [*0.5]) * (-1/N); x' = x_mean + g sqrt(1/N) z */
int sbm_predictor_step(const sbm_latent_shape* ls, const sbm_sde* sde, const float* x, const float* score,
                       const float* t, const float* noise, float* x_out, float* x_mean_out, int32_t probability_flow,
                       const sbm_rng* rng, const sbm_impute* impute, void* stream);
/* reverse-diffusion (ancestral) predictor: (f, G) = sde.discretize(x, t), rev_f = f - G^2 s [*0.5], x_mean = x - rev_f,
 * x' = x_mean + G z  (sde_helper2.py:236-253 base rule = subVPSDE, :373-381 VPSDE/DDPM, :465-473 VESDE/SMLD, :319-324
 * RSDE.discretize).  table = device copy of sde.discrete_betas (VPSDE) / sde.discrete_sigmas (VESDE), NULL for subVPSDE.
 * Same kernel, traffic (12 B / element) and epilogues as sbm_predictor_step. */
int sbm_rd_predictor_step(const sbm_latent_shape* ls, const sbm_sde* sde, const float* x, const float* score,
                          const float* t, const float* table, const float* noise, float* x_out, float* x_mean_out,
                          int32_t probability_flow, const sbm_rng* rng, const sbm_impute* impute, void* stream);
/* corrector, sde_helper2.py:96-98: acc2[0] += sum_b ||grad_b||, acc2[1] += sum_b ||noise_b||.  acc2 is a buffer of
 * THREE doubles, zero before the first call (the third is a completion ticket used by reset_acc below); multi-GPU
 * exact mode all-reduces acc2[0..1] between the two calls.  noise = injected buffer; else rng = Philox stream; with
 * BOTH NULL only acc2[0] is accumulated (4 B / element, memory-bound) and acc2[1] comes from sbm_noise_norm */
int sbm_corrector_norms(const sbm_latent_shape* ls, const float* grad, const float* noise, const sbm_rng* rng,
                        double* acc2, void* stream);
/* acc2[1] += sum_b ||noise_b|| of the Philox draw `rng` (sde_helper2.py:96, 98).  Touches no latent memory: the draw is
 * a function of (seed, draw id, element index), so the samplers run this on a side stream beside the score-net
 * forward and the norms kernel proper never regenerates the stream */
int sbm_noise_norm(const sbm_latent_shape* ls, const sbm_rng* rng, double* acc2, void* stream);
/* corrector, sde_helper2.py:56-60,99-101: step = (snr * mean||noise|| / mean||grad||)^2 * 2 * alpha[t];
 * x_mean = x + step*grad; x' = x_mean + sqrt(2 step) * noise.  alphas = device copy of sde.alphas (NULL: alpha = 1) */
int sbm_corrector_update(const sbm_latent_shape* ls, const sbm_sde* sde, const float* x, const float* grad,
                         const float* t, const float* noise, double* acc2, const float* alphas, float* x_out,
                         float* x_mean_out, float target_snr, int64_t global_batch, const sbm_rng* rng,
                         const sbm_impute* impute, int32_t reset_acc /* 1: zero acc2 once every block has read it */,
                         void* stream);
/* legacy annealed-Langevin evaluators (eval_lat_celeba_hq_all.py:268-275, fid_upd10.py:279-290): for every modality
 * channel m whose bit in obs_mask is clear, x' = x + coef_a[m] * score + coef_b[m] * noise; observed channels are copied
 * through.  coef_a / coef_b: HOST arrays of ls->mods floats (the per-level step sizes, folded on the host). */
int sbm_langevin_axpy_step(const sbm_latent_shape* ls, const float* x, const float* score, const float* noise,
                           const float* coef_a, const float* coef_b, uint32_t obs_mask, float* x_out,
                           const sbm_rng* rng, void* stream);
/* classifier / EBM guidance (sde_helper2.py:65-94, 283-312).  gather: new_x = cat(x[:, m1], x[:, m2]).view(B, 2*dd)
 * as bf16 rows of `ld` elements (the energy net's GEMM operand; padding zeroed).  apply: score[:, m1] -= cl_s *
 * grad[:, 0:dd], score[:, m2] -= cl_s * grad[:, dd:2dd] in place (grad = d mean(E) / d new_x, fp32 rows of ldg
 * elements); a negative m1 / m2 skips that half (train_poly_unet_cont.py:87).  The energy net itself (two hidden
 * layers) and its input gradient run on sbm_conv_igemm / sbm_time_embed / sbm_act_bwd. */
int sbm_guidance_gather(const sbm_latent_shape* ls, const float* x, int32_t m1, int32_t m2, void* out_bf16, int32_t ld,
                        void* stream);
int sbm_guidance_apply(const sbm_latent_shape* ls, float* score, const float* grad, int64_t ldg, int32_t m1, int32_t m2,
                       float cl_s, void* stream);
int sbm_impute_observed(const sbm_latent_shape* ls, const sbm_sde* sde, const float* x, float* x_out,
                        const sbm_impute* impute, void* stream);
/* loss_fn, sde_helper2.py:167-170: t = u*(T-eps)+eps; xt = mean(x0,t) + std(t)*z.  u,z injected or drawn (rng) */
int sbm_dsm_perturb(const sbm_latent_shape* ls, const sbm_sde* sde, const float* x0, const float* u, const float* z,
                    float* xt, float* z_out, float* t_out, float* std_out, float* g2_out, float eps,
                    const sbm_rng* rng, void* stream);
/* loss_fn, sde_helper2.py:173-185: *loss_acc += batch-mean loss (caller zeroes); dscore = d loss / d score */
int sbm_dsm_loss(const sbm_latent_shape* ls, const float* score, const float* z, const float* std, const float* g2,
                 float* dscore, double* loss_acc, int32_t likelihood_weighting, int32_t reduce_mean,
                 int64_t global_batch, void* stream);
/* nn.Dropout(p), training mode (unet_openai.py:265), IN PLACE on channels-last x[rows][C] (SBM_F32 / SBM_BF16, row
 * stride ld elements, ld % 8 == 0): x *= keep/(1-p), keep = Philox uniform of (seed, draw [+ *draw_dev], element) >= p.
 * The mask is never stored: the same call on the gradient of the output is the backward. */
int sbm_dropout(void* x, int32_t dtype, int64_t ld, int64_t rows, int32_t C, float p, const sbm_rng* rng, void* stream);
int sbm_scale_by_scalar(const float* in, const float* scalar_dev, float* out, int64_t n, void* stream);
int sbm_f64_to_f32(const double* in, float* out, int32_t n, void* stream);

/* ------------------------------------------------------------------ backward operators (DSM training step) */
/* out[c] += column sums of x[rows][C] (bias gradients); caller zeroes out */
int sbm_colsum(const void* x, int32_t dtype, int64_t ld, int64_t rows, int32_t C, float* out, void* stream);
/* GroupNorm backward. x = the tensor that was normalised (x = act(pre) when in_act != 0 and the kernel is given pre);
 * bst[B][G][2], dgamma[C], dbeta[C] are accumulated into (caller zeroes); dx (+ addend) -> out_f32 / out_bf16;
 * with in_act the result is additionally multiplied by act'(pre), i.e. it is the gradient w.r.t. pre. */
int sbm_groupnorm_bwd(const void* x, int32_t x_dtype, int64_t ldx, const void* dy, int32_t dy_dtype, int64_t lddy,
                      const double* stats, const float* gamma, float* bst, float* dgamma, float* dbeta,
                      const float* addend, int64_t ldadd, float* out_f32, int64_t ldo_f32, void* out_bf16,
                      int64_t ldo_bf16, int32_t B, int32_t HW, int32_t C, int32_t G, float eps, int32_t in_act,
                      const float* beta /* needed when out_act != 0 */,
                      int32_t out_act /* activation applied AFTER the norm (GroupNorm32 -> SiLU, unet_openai.py:252) */,
                      void* stream);
/* out[b][c] = sum over the pixels of sample b of x[b][p][c]: gradient of a per-sample row bias (unet_openai.py:303) */
int sbm_colsum_per_sample(const void* x, int32_t dtype, int64_t ld, int32_t B, int32_t HW, int32_t C, float* out,
                          int64_t ldo, void* stream);
/* backward of y = act(n * (1 + scale[b][c]) + shift[b][c]) (ResBlock(use_scale_shift_norm=True), unet_openai.py:296-300;
 * n = GroupNorm32 output, fp32 channels-last): dn = dy * act'(u) * (1 + scale); dscale[b][c] += sum_pix dy * act'(u) * n;
 * dshift[b][c] += sum_pix dy * act'(u)  (fp32 [B][ld_dmod], caller zeroes). */
int sbm_scale_shift_bwd(const float* n, int64_t ldn, const float* dy, int64_t lddy, const float* scale,
                        const float* shift, int64_t ld_mod, int32_t act, float* dn, int64_t lddn, float* dscale,
                        float* dshift, int64_t ld_dmod, int32_t B, int32_t HW, int32_t C, void* stream);
/* backward of sbm_upsample_nearest2x: out[b,i,j,c] = sum of the 2x2 block of dy (fp32 [B,2H,2W,lddy]) */
int sbm_upsample_nearest2x_bwd(const float* dy, int64_t lddy, float* out, int64_t ldo, void* out_bf16, int64_t ldb,
                               int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
/* depthwise 7x7: dw[C][49] += , db[C] += , dcond[b][c] = sum_p dy  (caller zeroes dw, db) */
int sbm_dwconv7_wgrad(const float* x, int64_t ldx, const float* dy, int64_t lddy, float* dw, float* db, float* dcond,
                      int64_t ldc, int32_t B, int32_t H, int32_t W, int32_t C, void* stream);
int sbm_linear_attn_bwd(const float* qkv, int64_t ldq, const float* dout, int64_t ldd, void* dqkv_bf16, int64_t ldg,
                        int32_t B, int32_t n, int32_t heads, float scale, void* stream);
int sbm_softmax_attn_bwd(const float* qkv, int64_t ldq, const float* dout, int64_t ldd, void* dqkv_bf16, int64_t ldg,
                         int32_t B, int32_t n, int32_t heads, int32_t dh, int32_t q_off, int32_t k_off, int32_t v_off,
                         int32_t head_stride, float scale, void* stream);
/* out = dy * act'(pre) */
int sbm_act_bwd(const float* dy, int64_t lddy, const void* pre, int32_t pre_dtype, int64_t ldp, float* out_f32,
                int64_t ldo, void* out_bf16, int64_t ldb, int64_t rows, int32_t C, int32_t act, void* stream);
int sbm_nchw_to_nhwc(const float* x, void* out_bf16, int64_t ldb, float* out_f32, int64_t ldf, int32_t B, int32_t C,
                     int32_t HW, void* stream);
/* out = a + b (b may be NULL: plain copy / bf16 cast); rows x C elements with row strides */
int sbm_add(const float* a, int64_t lda, const float* b, int64_t ldb, float* out, int64_t ldo, void* out_bf16,
            int64_t ldh, int64_t rows, int32_t C, void* stream);
/* torch.optim.Adam step over many tensors in ONE launch (train_lat_celebhq_unet_cont2.py:100, :477).
 * tensors_dev: device array of descriptors; chunks_dev: device array of (tensor index, chunk index) pairs. */
typedef struct sbm_adam_tensor { float* param; const float* grad; float* exp_avg; float* exp_avg_sq; int64_t n; } sbm_adam_tensor;
int sbm_adam_step(const sbm_adam_tensor* tensors_dev, const int32_t* chunks_dev, int32_t n_chunks,
                  int32_t chunk_elems, float lr, float beta1, float beta2, float eps, int32_t step, float grad_scale,
                  const int32_t* step_dev /* optional DEVICE counter of completed steps: overrides `step` (CUDA graphs) */,
                  void* stream);
/* utils.py:79-90 update_ema over many tensors in ONE launch: ema = ema*decay + src*(1-decay)
 * (train_lat_celebhq_unet_cont2_cond.py:129, 672-674). */
typedef struct sbm_ema_tensor { float* ema; const float* src; int64_t n; } sbm_ema_tensor;
int sbm_ema_step(const sbm_ema_tensor* tensors_dev, const int32_t* chunks_dev, int32_t n_chunks, int32_t chunk_elems,
                 float decay, void* stream);
/* per-step device state of a graph-replayed TRAINING step: *step_dev += 1 (Adam bias correction), *draw_dev += draw_inc
 * (Philox draw id of sbm_dsm_perturb, which consumes 2 per step).  One 1-thread kernel. */
int sbm_train_tick(int32_t* step_dev, uint64_t* draw_dev, uint64_t draw_inc, void* stream);

/* ------------------------------------------------------------------ residual autoencoders */
/* RBlock tail `self.sf(x + xhat)` + `down_pool` / `up_pool` (h_vae_model_copy.py:26-39), ResEncoder.ch_enc's
 * LeakyReLU + AvgPool2d(2) (:48-50) and, with slope 0, the ReLU after z_lin (:128): y = LeakyReLU_slope(x) then
 * mode 0 nothing | 1 average pooling by `rate` | 2 nearest up-sampling by `rate`.  x: channels-last [B,H,W,ldx]
 * (SBM_F32 / SBM_BF16); out_bf16: channels-last rows of ldo elements (channel padding zeroed) and / or, mode 0 only,
 * out_nchw_f32 [B][C][H][W]. */
int sbm_lrelu_resample(const void* x, int32_t in_dtype, int64_t ldx, void* out_bf16, int64_t ldo, float* out_nchw_f32,
                       int32_t B, int32_t H, int32_t W, int32_t C, float slope, int32_t mode, int32_t rate,
                       void* stream);
/* superset for the CelebA-HQ variants RBlockN / ResDecoderN (h_vae_model_copy.py:347-428): act 0 = LeakyReLU(slope),
 * 1 = exact GELU; mode 3 = BILINEAR up-sampling by `rate` (nn.Upsample(mode='bilinear'), align_corners=False). */
int sbm_act_resample(const void* x, int32_t in_dtype, int64_t ldx, void* out_bf16, int64_t ldo, float* out_nchw_f32,
                     int32_t B, int32_t H, int32_t W, int32_t C, int32_t act, float slope, int32_t mode, int32_t rate,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SBMAE_B200_H_ */
