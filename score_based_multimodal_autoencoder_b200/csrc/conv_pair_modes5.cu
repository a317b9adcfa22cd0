// Statically compiled epilogue modes [16, 17] of the CTA-pair convolution kernel (see conv_pair.cuh: SBM_EPI_MODES).
#include "conv_pair.cuh"

namespace sbm {

int launch_pair_static_g5(int mode_idx, int bn, const CUtensorMap& tmA, const CUtensorMap& tmB, const EpiMaps& em,
                            const ConvKernelParams& p, int m_tiles, int n_tiles, int nphase, cudaStream_t stream) {
  switch (mode_idx) {
#define SBM_EPI_PICK(idx, mode) SBM_EPI_PICK_##idx(idx, mode)
#define SBM_EPI_PICK_0(idx, mode) 
#define SBM_EPI_PICK_1(idx, mode) 
#define SBM_EPI_PICK_2(idx, mode) 
#define SBM_EPI_PICK_3(idx, mode) 
#define SBM_EPI_PICK_4(idx, mode) 
#define SBM_EPI_PICK_5(idx, mode) 
#define SBM_EPI_PICK_6(idx, mode) 
#define SBM_EPI_PICK_7(idx, mode) 
#define SBM_EPI_PICK_8(idx, mode) 
#define SBM_EPI_PICK_9(idx, mode) 
#define SBM_EPI_PICK_10(idx, mode) 
#define SBM_EPI_PICK_11(idx, mode) 
#define SBM_EPI_PICK_12(idx, mode) 
#define SBM_EPI_PICK_13(idx, mode) 
#define SBM_EPI_PICK_14(idx, mode) 
#define SBM_EPI_PICK_15(idx, mode) 
#define SBM_EPI_PICK_16(idx, mode) SBM_EPI_LAUNCH_CASE(idx, mode)
#define SBM_EPI_PICK_17(idx, mode) SBM_EPI_LAUNCH_CASE(idx, mode)
    SBM_EPI_MODES(SBM_EPI_PICK)
#undef SBM_EPI_PICK
    default: return -1;
  }
}

}  // namespace sbm
