"""CPU tests: the LitEma mirror (prediff_b200/ema.py) against the unmodified reference class (tests/golden/ema.npz), and
LatentDiffusion.ema_scope (latent_diffusion.py:280-293) swapping the EMA weights into the CUDA UNet mirror."""
import os

import numpy as np
import torch

from prediff_b200 import weights as Wt
from prediff_b200.diffusion import LatentDiffusion
from prediff_b200.ema import LitEma
from prediff_b200.unet import CuboidTransformerUNet
from tests.golden.gen_golden import ema_drift, ema_model

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ema.npz"))


def test_lit_ema_matches_reference_bit_exactly():
    m = ema_model()
    ema = LitEma(m, decay=0.9)
    assert sorted(k for k, _ in ema.named_buffers()) == list(G["keys"])   # reference buffer names (dots removed)
    assert "0bias" not in dict(ema.named_buffers())                        # frozen parameters have no shadow
    for step in range(12):
        ema_drift(m, step)
        ema(m)
        if step + 1 in (1, 2, 12):
            for k, v in ema.named_buffers():
                assert np.array_equal(v.numpy(), G[f"s{step + 1}_{k}"]), (step, k)
    before = [p.clone() for p in m.parameters()]
    ema.store(m.parameters())
    ema.copy_to(m)
    assert np.array_equal(m[0].weight.detach().numpy(), G["swapped_0weight"])
    assert np.array_equal(m[0].bias.detach().numpy(), G["swapped_0bias_frozen"])
    ema.restore(m.parameters())
    assert all(torch.equal(a, b) for a, b in zip(before, m.parameters()))


def test_ema_scope_swaps_weights_into_the_cuda_unet_mirror():
    cfg = Wt.TINY_UNET
    unet = CuboidTransformerUNet([cfg.t_in, cfg.h, cfg.w, cfg.c], [cfg.t_out, cfg.h, cfg.w, cfg.c], base_units=64, depth=[1, 1])
    ld = LatentDiffusion(torch_nn_module=unet, use_ema=True)
    sd = ld.state_dict()
    assert "model_ema.decay" in sd and "model_ema.num_updates" in sd and "model_ema.final_projweight" in sd
    name = "down_self_blocks.0.0.attn_l.0.qkv.weight"
    p = dict(unet.named_parameters())[name]
    with torch.no_grad():
        p.add_(1.0)                      # "training" moved the live weights away from the shadow
    live = p.detach().clone()
    shadow = getattr(ld.model_ema, name.replace(".", "")).clone()
    assert not torch.equal(live, shadow)
    unet._dirty = False
    with ld.ema_scope():
        assert torch.equal(p.detach(), shadow) and unet._dirty    # EMA weights in place, CUDA side told to re-read them
        unet._dirty = False
    assert torch.equal(p.detach(), live) and unet._dirty
    ld.on_train_batch_end()
    assert int(ld.model_ema.num_updates) == 1
    # use_ema=False (the default of the mirror): the scope is a no-op and nothing is registered
    ld0 = LatentDiffusion(torch_nn_module=unet)
    assert not hasattr(ld0, "model_ema")
    with ld0.ema_scope("ctx"):
        pass


def test_latent_diffusion_state_dict_keys_equal_the_reference():
    """state_dict() of the mirror (UNet + VAE + schedule buffers + logvar + model_ema.*) has exactly the keys and shapes
    of the unmodified reference LatentDiffusion(use_ema=True) - a checkpoint of the reference module loads strictly."""
    from prediff_b200.vae import AutoencoderKL
    ucfg, vcfg = Wt.TINY_UNET, Wt.TINY_VAE
    unet = CuboidTransformerUNet([ucfg.t_in, ucfg.h, ucfg.w, ucfg.c], [ucfg.t_out, ucfg.h, ucfg.w, ucfg.c], base_units=64, depth=[1, 1])
    vae = AutoencoderKL(block_out_channels=vcfg.block_out_channels, layers_per_block=vcfg.layers_per_block,
                        latent_channels=vcfg.latent_channels, sample_size=(vcfg.h, vcfg.w))
    ld = LatentDiffusion(torch_nn_module=unet, use_ema=True, first_stage_model=vae, cond_stage_model="__is_first_stage__")
    mine = {k: ",".join(str(d) for d in v.shape) for k, v in ld.state_dict().items()}
    ref = dict(e.split("|") for e in G["ldm_state_dict_keys"])
    assert set(mine) == set(ref), (sorted(set(ref) - set(mine))[:5], sorted(set(mine) - set(ref))[:5])
    assert mine == ref
    ld.load_state_dict(ld.state_dict(), strict=True)


def test_step_helper_methods_match_reference_bit_exactly():
    """predict_start_from_noise / q_posterior / p_mean_variance / aligned_mean (latent_diffusion.py:553-596) of the
    mirror vs the unmodified reference (tests/golden/helpers.npz), with a closed-form denoiser on CPU tensors."""
    from tests.golden.gen_golden import DummyEps, inp
    H = np.load(os.path.join(os.path.dirname(__file__), "golden", "helpers.npz"))
    ld = LatentDiffusion(torch_nn_module=DummyEps())
    zt, zc, noise = inp(61, 3, 6, 4, 4, 8), inp(62, 3, 7, 4, 4, 8), inp(63, 3, 6, 4, 4, 8)
    t = torch.tensor([0, 431, 999], dtype=torch.long)

    def same(a, key):
        assert np.array_equal(a.numpy(), H[key]), key

    same(ld.predict_start_from_noise(zt, t, noise), "x0")
    for a, k in zip(ld.q_posterior(noise, zt, t), ("qp_mean", "qp_var", "qp_logvar")):
        same(a, k)
    for a, k in zip(ld.p_mean_variance(zt, zc, t, clip_denoised=False, return_x0=True), ("pm_mean", "pm_var", "pm_logvar", "pm_x0")):
        same(a, k)
    same(ld.p_mean_variance(zt, zc, t, clip_denoised=True)[0], "pm_mean_clipped")
    ld.set_alignment(lambda zt, t, zc=None, y=None, **kw: 0.2 * zt + kw["shift"])
    same(ld.aligned_mean(zt, t, zc, None, torch.from_numpy(H["pm_mean"]), torch.from_numpy(H["pm_logvar"]), shift=0.05), "aligned")
    assert ld.get_batch_data_shape(5) == (5, 6, 128, 128, 1) and ld.einops_spatial_layout == "(N T) C H W"
    assert torch.equal(ld.get_first_stage_encoding(zt), zt)
