"""GPU parity: AutoencoderKL CUDA path vs golden outputs of the unmodified reference and vs the CPU oracle.
Tolerances (bf16 tensor-core operands, fp32 accumulate / residual stream; measured 9e-3 on B200): rel-RMS <= 1.4e-2,
max <= 4e-2 of abs-max."""
import os

import numpy as np
import pytest
import torch

from oracle import prediff_oracle as O
from prediff_b200 import weights as Wt
from prediff_b200.vae import AutoencoderKL
from tests.golden.gen_golden import VAE_SEED, inp
from tests.test_unet_gpu import errs

REL_RMS_TOL, MAX_TOL = 1.4e-2, 4e-2

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def make_vae(cfg, max_frames=16):
    m = AutoencoderKL(in_channels=cfg.in_channels, out_channels=cfg.out_channels,
                      block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block,
                      latent_channels=cfg.latent_channels, norm_num_groups=cfg.norm_num_groups,
                      sample_size=(cfg.h, cfg.w), max_frames=max_frames)
    sd = O.to_torch_sd(Wt.seeded_state_dict(Wt.vae_param_spec(cfg), VAE_SEED))
    m.load_state_dict(sd, strict=True)
    return m.eval(), sd


@pytest.mark.parametrize("tag,cfg,N", [("tiny", Wt.TINY_VAE, 2), ("full", Wt.VAEConfig(), 1)])
def test_vae_vs_reference_golden(tag, cfg, N):
    """BASELINE config 1 (encode + decode of a 128x128 frame), here on the GPU path."""
    m, _ = make_vae(cfg)
    g = np.load(os.path.join(G, f"vae_{tag}.npz"))
    x = inp(4321, N, 1, cfg.h, cfg.w, uniform=True).cuda()
    post = m.encode(x)
    assert tuple(post.parameters.shape) == (N, 2 * cfg.latent_channels, cfg.h // 8, cfg.w // 8)
    r1, m1 = errs(post.parameters, g["moments"])
    # decode the reference's own latent so the decoder is checked in isolation
    z_ref = torch.as_tensor(g["moments"][:, :cfg.latent_channels]).cuda()
    dec = m.decode(z_ref)
    assert tuple(dec.shape) == (N, 1, cfg.h, cfg.w)
    r2, m2 = errs(dec, g["dec"])
    print(f"vae {tag}: encode rel_rms={r1:.3e} max={m1:.3e}; decode rel_rms={r2:.3e} max={m2:.3e}")
    assert r1 < REL_RMS_TOL and m1 < MAX_TOL and r2 < REL_RMS_TOL and m2 < MAX_TOL
    # mode() is the mean half (distributions.py:70-71)
    assert torch.equal(post.mode(), post.parameters[:, :cfg.latent_channels])


def test_vae_tiny_vs_oracle_ragged_batches():
    cfg = Wt.TINY_VAE
    m, sd = make_vae(cfg, max_frames=3)
    x = inp(99, 7, 1, cfg.h, cfg.w, uniform=True)  # 7 frames through max_frames=3 chunks (3 + 3 + 1)
    with torch.no_grad():
        mom = O.vae_encode_moments(sd, cfg, x)
        dec = O.vae_decode(sd, cfg, mom[:, :cfg.latent_channels])
    got = m.encode_moments(x.cuda())
    r1, m1 = errs(got, mom)
    d = m.decode(mom[:, :cfg.latent_channels].cuda())
    r2, m2 = errs(d, dec)
    print(f"vae tiny x7: encode rel_rms={r1:.3e} decode rel_rms={r2:.3e}")
    assert r1 < REL_RMS_TOL and m1 < MAX_TOL and r2 < REL_RMS_TOL and m2 < MAX_TOL
    # frames are independent: frame 5 alone equals frame 5 of the batch
    one = m.encode_moments(x[5:6].cuda())
    assert torch.equal(one[0], got[5])
