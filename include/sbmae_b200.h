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
 *    ownership transfer, no hidden global state except immutable caches;
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
       SBM_CONVT_4X4_S2 = 2 /* ConvTranspose2d 4x4, stride 2, padding 1             */ };
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
  int32_t reserved;
  double* stats;            /* [batch][2] += (sum, sum of squares) of the written values, or NULL */
  void* out2;               /* optional second copy of the output in bf16 (pixel stride ldo2), or NULL */
  int64_t ldo2;
} sbm_conv_args;

int sbm_conv_igemm(const sbm_conv_args* a, void* stream);

/* fp32 weights -> bf16 [taps][rows][cols_pad]; src element (tap,row,col) at
 * w[tap*s_tap + row*s_row + col*s_col]; optional per-column scale (GroupNorm gamma folding). */
int sbm_pack_weight_bf16(const float* w, void* dst, int32_t taps, int32_t rows, int32_t cols, int32_t cols_pad,
                         int64_t s_tap, int64_t s_row, int64_t s_col, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SBMAE_B200_H_ */
