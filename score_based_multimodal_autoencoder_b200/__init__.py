"""B200-native implementation of SBM-AE's latent score-model hot path.

Drop-in modules (same names as the reference's files):
    score_based_multimodal_autoencoder_b200.sde_helper2   VPSDE/subVPSDE/VESDE, em_predictor, corrector,
                                                          uncond_sampler, loss_fn (+ cond_sampler, pc_sampler)
    score_based_multimodal_autoencoder_b200.unet_model    Unet (ConvNeXt score net)
    score_based_multimodal_autoencoder_b200.unet_openai   UNetModel (guided-diffusion score net)
All arithmetic of the path runs in libsbmae_b200.so (hand-written sm_100a CUDA, C ABI in include/sbmae_b200.h).
"""
__version__ = "0.1.0"
