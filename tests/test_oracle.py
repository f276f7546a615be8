"""CPU tests: pins the oracle (oracle/prediff_oracle.py) against the golden fixtures generated from the
unmodified reference (tests/golden/gen_golden.py) and against the schedule known-answers of SURVEY.md section 4."""
import os

import numpy as np
import pytest
import torch

from oracle import prediff_oracle as O
from prediff_b200 import weights as Wt
from tests.golden.gen_golden import UNET_SEED, VAE_SEED, inp

G = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return np.load(os.path.join(G, name + ".npz"))


def maxrel(a, b):
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.fixture(scope="module")
def tiny_unet_sd():
    return O.to_torch_sd(Wt.seeded_state_dict(Wt.unet_param_spec(Wt.TINY_UNET), UNET_SEED))


def test_schedule_known_answers():
    s = O.make_schedule()
    ac = s["alphas_cumprod"].numpy()
    # SURVEY.md section 4 item 1 (probed from the reference's diffusion/utils.py + register_schedule)
    np.testing.assert_allclose(ac[[0, 1, 500, 981, 999]],
                               [0.9999, 0.999797362, 0.331274580, 1.04898745e-3, 7.33412460e-4], rtol=2e-7)
    assert abs(s["posterior_log_variance_clipped"][0].item() - (-46.0517019)) < 1e-5
    assert abs(s["sqrt_recip_alphas_cumprod"][999].item() - 36.9254552) < 1e-5
    ts = O.ddim_timesteps(50, 1000)
    assert ts[0] == 1 and ts[1] == 21 and ts[-1] == 981 and len(ts) == 50


def test_schedule_vs_reference_buffers():
    g = gold("schedule")
    s = O.make_schedule()
    for k, v in s.items():
        assert np.array_equal(v.numpy(), g[k]), k  # bit-exact fp32 buffers
    ts = O.ddim_timesteps(50, 1000)
    assert np.array_equal(ts, g["ddim50_timesteps"])
    for eta in (0.0, 1.0):
        sig, a, ap = O.ddim_params(s["alphas_cumprod"].numpy(), ts, eta)
        np.testing.assert_array_equal(sig, g[f"ddim50_sigmas_eta{int(eta)}"])
    np.testing.assert_array_equal(a, g["ddim50_alphas"])
    np.testing.assert_array_equal(ap, g["ddim50_alphas_prev"])
    e = O.timestep_embedding(torch.tensor([0, 1, 500, 981, 999]), 256)
    assert np.array_equal(e.numpy(), g["timestep_embedding_256"])


def test_unet_tiny_vs_reference(tiny_unet_sd):
    cfg = Wt.TINY_UNET
    g = gold("unet_tiny")
    x = inp(1234, 2, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cond = inp(1235, 2, cfg.t_in, cfg.h, cfg.w, cfg.c)
    with torch.no_grad():
        out = O.unet_forward(tiny_unet_sd, cfg, x, torch.as_tensor(g["t"]), cond)
    assert maxrel(out, g["out"]) < 2e-5


def test_vae_tiny_vs_reference():
    cfg = Wt.TINY_VAE
    sd = O.to_torch_sd(Wt.seeded_state_dict(Wt.vae_param_spec(cfg), VAE_SEED))
    g = gold("vae_tiny")
    x = inp(4321, 2, 1, cfg.h, cfg.w, uniform=True)
    with torch.no_grad():
        mom = O.vae_encode_moments(sd, cfg, x)
        dec = O.vae_decode(sd, cfg, mom[:, :cfg.latent_channels])
    assert maxrel(mom, g["moments"]) < 2e-5
    assert maxrel(dec, g["dec"]) < 2e-5


def test_ddpm_loop_and_step_tiny_vs_reference(tiny_unet_sd):
    cfg = Wt.TINY_UNET
    g = gold("loop_tiny")
    sched = O.make_schedule()
    B, n_steps = 2, 4
    zT = inp(777, B, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cond = inp(778, B, cfg.t_in, cfg.h, cfg.w, cfg.c)
    noise = inp(779, n_steps, B, cfg.t_out, cfg.h, cfg.w, cfg.c)
    with torch.no_grad():
        z0 = O.sample_loop_ddpm(tiny_unet_sd, cfg, sched, zT.clone(), cond, noise, n_steps)
        eps = O.unet_forward(tiny_unet_sd, cfg, zT, torch.full((B,), 900), cond)
        zs = O.p_sample_ddpm(sched, eps, zT, 900, noise[0])
    assert maxrel(z0, g["z0"]) < 5e-5
    assert maxrel(zs, g["z_step900"]) < 2e-5


def test_ddpm_loop_shorten_cond_schedule_vs_reference(tiny_unet_sd):
    """num_timesteps_cond = 4: the reference re-noises the context before every step (latent_diffusion.py:295-299,
    665-667); golden from the unmodified p_sample_loop with both RNG streams injected."""
    cfg = Wt.TINY_UNET
    g = gold("loop_shorten")
    sched = O.make_schedule()
    B, n_steps = 2, 4
    zT = inp(777, B, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cond = inp(778, B, cfg.t_in, cfg.h, cfg.w, cfg.c)
    noise = inp(779, n_steps, B, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cnoise = inp(781, n_steps, B, cfg.t_in, cfg.h, cfg.w, cfg.c)
    ids = O.cond_schedule_ids(4)
    assert torch.equal(ids, torch.as_tensor(g["cond_ids"]))
    with torch.no_grad():
        z0 = O.sample_loop_ddpm(tiny_unet_sd, cfg, sched, zT.clone(), cond, noise, n_steps, cond_ids=ids, cond_noise=cnoise)
    assert maxrel(z0, g["z0"]) < 5e-5


def test_sample_end_to_end_tiny_vs_reference(tiny_unet_sd):
    """LatentDiffusion.sample(): encode context (mode) -> DDPM loop -> decode (latent_diffusion.py:686-724)."""
    cfg = Wt.TINY_UNET
    vcfg = Wt.TINY_VAE
    vsd = O.to_torch_sd(Wt.seeded_state_dict(Wt.vae_param_spec(vcfg), VAE_SEED))
    g = gold("loop_tiny")
    sched = O.make_schedule()
    B, n_steps = 2, 4
    zT = inp(777, B, cfg.t_out, cfg.h, cfg.w, cfg.c)
    noise = inp(779, n_steps, B, cfg.t_out, cfg.h, cfg.w, cfg.c)
    y = inp(780, B, cfg.t_in, vcfg.h, vcfg.w, 1, uniform=True)
    with torch.no_grad():
        frames = y.permute(0, 1, 4, 2, 3).reshape(B * cfg.t_in, 1, vcfg.h, vcfg.w)
        zc = O.vae_encode_mode(vsd, vcfg, frames)  # (B*T, C, h, w)
        zc = zc.reshape(B, cfg.t_in, cfg.c, cfg.h, cfg.w).permute(0, 1, 3, 4, 2)
        assert maxrel(zc, g["sample_zc"]) < 2e-5
        z0 = O.sample_loop_ddpm(tiny_unet_sd, cfg, sched, zT.clone(), zc, noise, n_steps)
        dec = O.vae_decode(vsd, vcfg, z0.permute(0, 1, 4, 2, 3).reshape(B * cfg.t_out, cfg.c, cfg.h, cfg.w))
        dec = dec.reshape(B, cfg.t_out, 1, vcfg.h, vcfg.w).permute(0, 1, 3, 4, 2)
    assert maxrel(dec, g["sample_dec"]) < 1e-4


def test_ddim_tiny_vs_reference(tiny_unet_sd):
    cfg = Wt.TINY_UNET
    g = gold("ddim_tiny")
    sched = O.make_schedule()
    z = inp(4242, 2, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cond = inp(4243, 2, cfg.t_in, cfg.h, cfg.w, cfg.c)
    with torch.no_grad():
        z0 = O.sample_loop_ddim(tiny_unet_sd, cfg, sched, z, cond, 50)
    assert maxrel(z0, g["z0"]) < 2e-4


@pytest.mark.slow
def test_unet_full_vs_reference():
    cfg = Wt.UNetConfig()
    sd = O.to_torch_sd(Wt.seeded_state_dict(Wt.unet_param_spec(cfg), UNET_SEED))
    g = gold("unet_full")
    x = inp(1234, 1, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cond = inp(1235, 1, cfg.t_in, cfg.h, cfg.w, cfg.c)
    with torch.no_grad():
        out = O.unet_forward(sd, cfg, x, torch.as_tensor(g["t"]), cond)
    assert maxrel(out, g["out"]) < 5e-5


def test_vae_full_vs_reference():
    cfg = Wt.VAEConfig()
    sd = O.to_torch_sd(Wt.seeded_state_dict(Wt.vae_param_spec(cfg), VAE_SEED))
    g = gold("vae_full")
    x = inp(4321, 1, 1, cfg.h, cfg.w, uniform=True)
    with torch.no_grad():
        mom = O.vae_encode_moments(sd, cfg, x)
        dec = O.vae_decode(sd, cfg, mom[:, :cfg.latent_channels])
    assert maxrel(mom, g["moments"]) < 5e-5
    assert maxrel(dec, g["dec"]) < 5e-5


def test_knowledge_alignment_vs_reference(tiny_unet_sd):
    """U(z_t, t), the guidance gradient (get_mean_shift) and one aligned DDPM step vs the reference's own."""
    from tests.golden.gen_golden import KA_SEED
    cfg = Wt.KAConfig()
    sd = O.to_torch_sd(Wt.seeded_state_dict(Wt.ka_param_spec(cfg), KA_SEED))
    g = gold("ka_full")
    zt = inp(5151, 4, cfg.t, cfg.h, cfg.w, cfg.c)
    t = torch.as_tensor(g["t"])
    target = torch.full((4, 1), 0.3)
    with torch.no_grad():
        pred = O.ka_forward(sd, cfg, zt, t)
    assert maxrel(pred, g["pred"]) < 5e-5
    grad = O.ka_mean_shift(sd, cfg, zt, t, target, cfg.guide_scale)
    assert maxrel(grad, g["grad"]) < 2e-4
    # aligned p_sample (latent_diffusion.py:592-631) with the tiny UNet
    ucfg = Wt.TINY_UNET
    sched = O.make_schedule()
    zT = inp(777, 2, ucfg.t_out, ucfg.h, ucfg.w, ucfg.c)
    cond = inp(778, 2, ucfg.t_in, ucfg.h, ucfg.w, ucfg.c)
    noise = inp(779, 4, 2, ucfg.t_out, ucfg.h, ucfg.w, ucfg.c)
    ts = torch.full((2,), 900)
    with torch.no_grad():
        eps = O.unet_forward(tiny_unet_sd, ucfg, zT, ts, cond)
    guide = O.ka_mean_shift(sd, cfg, zT, ts, torch.full((2, 1), 0.3), cfg.guide_scale)
    z = O.p_sample_ddpm(sched, eps, zT, 900, noise[0], guide=guide)
    assert maxrel(z, g["z_aligned_step900"]) < 5e-5


def test_forward_loss_vs_reference(tiny_unet_sd):
    """p_losses / q_sample / lvlb_weights of the oracle vs the unmodified reference LatentDiffusion (losses.npz)."""
    from tests.golden.gen_golden import LOSS_CASES
    g = gold("losses")
    cfg = Wt.TINY_UNET
    sched = O.make_schedule()
    assert np.array_equal(O.lvlb_weights(sched).numpy(), g["lvlb_weights"])
    z, zc, noise = inp(881, 3, cfg.t_out, cfg.h, cfg.w, cfg.c), inp(882, 3, cfg.t_in, cfg.h, cfg.w, cfg.c), \
        inp(883, 3, cfg.t_out, cfg.h, cfg.w, cfg.c)
    t = torch.as_tensor(g["t"])
    assert np.array_equal(O.q_sample(sched, z, t, noise)[:, 0, 0, 0, :8].numpy(), g["x_noisy_probe"])
    for tag, kw in LOSS_CASES:
        with torch.no_grad():
            r = O.p_losses(tiny_unet_sd, cfg, sched, z, zc, t, noise, **kw)
        for k in ("loss_simple", "loss_vlb", "loss"):
            want = float(g[f"{tag}_val_{k}"])
            assert abs(float(r[k]) - want) <= 2e-5 * abs(want), (tag, k, float(r[k]), want)
        assert abs(float(r["loss"]) - float(g[f"{tag}_loss"])) <= 2e-5 * abs(float(g[f"{tag}_loss"]))


def test_mirror_loss_terms_vs_reference_incl_learned_logvar(tiny_unet_sd):
    """LatentDiffusion.p_losses of the mirror (tensor path, with the oracle UNet standing in as the denoiser module on
    CPU) vs the unmodified reference: fixed logvar, weighted l1 + vlb, and learn_logvar = True as in the shipped
    config (cfg.yaml:95) with per-timestep logvar values - dict keys, order and values."""
    from prediff_b200.diffusion import LatentDiffusion
    from tests.golden.gen_golden import LEARNED_LOGVAR_CASE, LOSS_CASES, learned_logvar_values
    g = gold("losses")
    cfg = Wt.TINY_UNET

    class OracleEps(torch.nn.Module):
        def forward(self, x, t, cond):
            return O.unet_forward(tiny_unet_sd, cfg, x, t, cond)

    z, zc, noise = inp(881, 3, cfg.t_out, cfg.h, cfg.w, cfg.c), inp(882, 3, cfg.t_in, cfg.h, cfg.w, cfg.c), \
        inp(883, 3, cfg.t_out, cfg.h, cfg.w, cfg.c)
    t = torch.as_tensor(g["t"])
    for tag, kw in LOSS_CASES + [LEARNED_LOGVAR_CASE]:
        ld = LatentDiffusion(torch_nn_module=OracleEps(), **kw).eval()
        if kw.get("learn_logvar"):
            assert isinstance(ld.logvar, torch.nn.Parameter) and "logvar" in ld.state_dict()
            ld.logvar.data.copy_(learned_logvar_values())
        loss, d = ld.p_losses(z, zc, t, noise=noise)
        want_keys = ["val/loss_simple"] + (["val/loss_gamma", "logvar"] if kw.get("learn_logvar") else []) + \
            ["val/loss_vlb", "val/loss"]
        assert list(d) == want_keys
        for k, v in d.items():
            want = float(g[f"{tag}_{k.replace('/', '_')}"])
            assert abs(float(v) - want) <= 2e-5 * max(abs(want), 1e-3), (tag, k, float(v), want)
        assert abs(float(loss) - float(g[f"{tag}_loss"])) <= 2e-5 * abs(float(g[f"{tag}_loss"]))
