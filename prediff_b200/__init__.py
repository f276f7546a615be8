"""prediff_b200: B200-native (sm_100a) latent-diffusion sampling path for PreDiff.

Mirrors the reference call surface:
  LatentDiffusion.sample / p_sample_loop, CuboidTransformerUNet.forward(x, t, cond), AutoencoderKL.encode / decode,
driving hand-written CUDA kernels through the C ABI in include/prediff_b200.h.
"""
__version__ = "0.1.0"
