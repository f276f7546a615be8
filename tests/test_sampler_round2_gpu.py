"""Round-2 GPU parity cases of the sampler (VERDICT r01 'weak' 2, 3, 14):
clip_denoised against the reference's p_sample / p_sample_loop, the full-size `sample()` (shipped frames -> encode ->
loop -> decode) against goldens of the unmodified reference, the ABI entries pd_sample_loop / pd_sample_step_ddpm, and
the cached CUDA graphs surviving (a) new caller addresses and (b) a weight refresh of the denoiser."""
import ctypes
import os

import numpy as np
import pytest
import torch

from prediff_b200 import _lib as L
from prediff_b200 import weights as Wt
from prediff_b200.diffusion import PD_MODE_DDIM, PD_MODE_DDPM, LatentDiffusion
from tests.golden.gen_golden import inp
from tests.test_unet_gpu import errs, make_unet
from tests.test_vae_gpu import make_vae

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "loop_extra.npz"))
CFG = Wt.TINY_UNET
# bf16 operands (default precision); see tests/test_tf32_gpu.py for the TF32-class bars
# latents: as tests/test_sampler_gpu.py; decoded full-size frames after encode -> 50 steps -> decode measure 1.6e-2 (bf16 VAE on both sides)
LOOP_RMS_TOL, LOOP_MAX_TOL = 1.2e-2, 2.5e-2
FRAME_RMS_TOL, FRAME_MAX_TOL = 2.5e-2, 4e-2


def _tiny_inputs(scale=1.0):
    zT = inp(777, 2, CFG.t_out, CFG.h, CFG.w, CFG.c).cuda() * scale
    cond = inp(778, 2, CFG.t_in, CFG.h, CFG.w, CFG.c).cuda()
    noise = inp(779, 4, 2, CFG.t_out, CFG.h, CFG.w, CFG.c).cuda()
    return zT, cond, noise


def test_shorten_cond_schedule_loop_vs_reference():
    """num_timesteps_cond = 4 (latent_diffusion.py:153-157, 295-299, 665-667): the context latents are re-noised before
    every ancestral step; golden of the unmodified p_sample_loop with both RNG streams injected."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "loop_shorten.npz"))
    unet, _ = make_unet(CFG)
    ldm = LatentDiffusion(torch_nn_module=unet, num_timesteps_cond=4)
    assert ldm.shorten_cond_schedule and np.array_equal(ldm.cond_ids.numpy(), g["cond_ids"])
    assert "cond_ids" in ldm.state_dict()
    zT, cond, noise = _tiny_inputs()
    cnoise = inp(781, 4, 2, CFG.t_in, CFG.h, CFG.w, CFG.c).cuda()
    z0 = ldm.p_sample_loop(cond=cond, shape=tuple(zT.shape), x_T=zT, timesteps=4, noise=noise, cond_noise=cnoise)
    r, m = errs(z0, g["z0"])
    plain = LatentDiffusion(torch_nn_module=unet).p_sample_loop(cond=cond, shape=tuple(zT.shape), x_T=zT, timesteps=4,
                                                                noise=noise)
    r_plain, _ = errs(plain, g["z0"])
    print(f"shorten_cond_schedule 4-step loop: rel_rms={r:.3e} max={m:.3e} (fixed context: {r_plain:.3e})")
    assert r < LOOP_RMS_TOL and m < LOOP_MAX_TOL and r_plain > 5 * r


def test_clip_denoised_step_and_loop_vs_reference():
    """latent_diffusion.py:580-581: z_recon.clamp_(-1, 1) inside p_mean_variance, here inside the fused update kernel."""
    unet, _ = make_unet(CFG)
    ldm = LatentDiffusion(torch_nn_module=unet)
    zT, cond, noise = _tiny_inputs()
    zh = zT * 0.5
    t = torch.full((2,), 100, device="cuda")
    out = ldm.p_sample(zt=zh, zc=cond, t=t, clip_denoised=True, noise=noise[0])
    r, m = errs(out, G["z_step100_clip"])
    # the same call without the clamp must differ: the golden really exercises the clamp (~26 % of the entries)
    plain = ldm.p_sample(zt=zh, zc=cond, t=t, clip_denoised=False, noise=noise[0])
    r_plain, _ = errs(plain, G["z_step100_clip"])
    print(f"clip_denoised p_sample: rel_rms={r:.3e} max={m:.3e} (without the clamp {r_plain:.3e})")
    assert r < LOOP_RMS_TOL and m < LOOP_MAX_TOL and r_plain > 5 * r
    # return_x0 route (tensor expressions on the CUDA UNet's eps) gives the clamped estimate itself
    _, x0 = ldm.p_sample(zt=zh, zc=cond, t=t, clip_denoised=True, noise=noise[0], return_x0=True)
    assert float(x0.abs().max()) <= 1.0
    rx, mx = errs(x0, G["z_step100_clip_x0"])
    assert rx < LOOP_RMS_TOL and mx < LOOP_MAX_TOL
    ldm_c = LatentDiffusion(torch_nn_module=unet, clip_denoised=True)
    z0 = ldm_c.p_sample_loop(cond=cond, shape=tuple(zT.shape), x_T=zT, timesteps=4, noise=noise)
    r, m = errs(z0, G["z0_clip"])
    print(f"clip_denoised 4-step loop: rel_rms={r:.3e} max={m:.3e}")
    assert r < LOOP_RMS_TOL and m < LOOP_MAX_TOL
    # host-driven loop (p_sample per step through pd_sample_step_ddpm) == device-resident loop, bit for bit
    img = zT
    for k, i in enumerate(reversed(range(4))):
        img = ldm_c.p_sample(zt=img, zc=cond, t=torch.full((2,), i, device="cuda"), clip_denoised=True, noise=noise[k])
    assert torch.equal(img, z0)


def test_abi_sample_loop_and_step_entries():
    """pd_sample_loop == pd_sample_loop_range over the whole schedule == n calls of pd_sample_step_ddpm."""
    unet, _ = make_unet(CFG)
    ldm = LatentDiffusion(torch_nn_module=unet)
    zT, cond, noise = _tiny_inputs()
    lib = L.lib()
    a, b, c = zT.clone(), zT.clone(), zT.clone()
    L.check(lib.pd_sample_loop(ldm._sampler, unet.handle, L.ptr(a), L.ptr(cond), L.ptr(noise), 2, PD_MODE_DDPM, 4,
                               ctypes.c_float(0.0), L.stream_ptr()))
    L.check(lib.pd_sample_loop_range(ldm._sampler, unet.handle, L.ptr(b), L.ptr(cond), L.ptr(noise), 2, PD_MODE_DDPM, 4,
                                     ctypes.c_float(0.0), 0, 4, L.stream_ptr()))
    for k, t in enumerate(reversed(range(4))):
        nk = noise[k].contiguous()
        L.check(lib.pd_sample_step_ddpm(ldm._sampler, unet.handle, L.ptr(c), L.ptr(cond), L.ptr(nk), 2, t, L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(a, c)
    r, m = errs(a, np.load(os.path.join(os.path.dirname(__file__), "golden", "loop_tiny.npz"))["z0"])
    assert r < LOOP_RMS_TOL and m < LOOP_MAX_TOL
    # DDIM through the same entry, deterministic (eta = 0, noise = NULL)
    d1, d2 = zT.clone(), zT.clone()
    for d in (d1, d2):
        L.check(lib.pd_sample_loop(ldm._sampler, unet.handle, L.ptr(d), L.ptr(cond), None, 2, PD_MODE_DDIM, 10,
                                   ctypes.c_float(0.0), L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(d1, d2) and torch.isfinite(d1).all()


def test_graph_survives_new_addresses_and_weight_refresh():
    """The cached loop graph runs on sampler-owned buffers (new caller tensors need no re-capture) and is keyed on the
    denoiser's weight generation: after a refresh (load_state_dict through the PARENT module, ema-style .data writes) the
    next loop must use the new weights - compared bit-exactly with a model that only ever saw them."""
    unet, sd = make_unet(CFG)
    ldm = LatentDiffusion(torch_nn_module=unet)
    zT, cond, _ = _tiny_inputs()
    shape = tuple(zT.shape)
    z_a = ldm.ddim_sample_loop(cond=cond, shape=shape, x_T=zT, ddim_steps=10)
    keep = [torch.empty_like(zT) for _ in range(3)]   # shifts the allocator: the next clone lands at a new address
    z_b = ldm.ddim_sample_loop(cond=cond.clone(), shape=shape, x_T=zT.clone(), ddim_steps=10)
    assert torch.equal(z_a, z_b)
    del keep
    # new weights through the parent's load_state_dict (nn.Module recursion never calls the child's override)
    full = ldm.state_dict()
    g = torch.Generator().manual_seed(99)
    for k in full:
        if k.startswith("torch_nn_module.") and full[k].dtype == torch.float32 and full[k].dim() >= 2:
            full[k] = full[k] * (1.0 + 0.05 * torch.randn(full[k].shape, generator=g))
    ldm.load_state_dict(full)
    z_c = ldm.ddim_sample_loop(cond=cond, shape=shape, x_T=zT, ddim_steps=10)
    assert not torch.equal(z_c, z_a)
    fresh, _ = make_unet(CFG)
    fresh.load_state_dict({k[len("torch_nn_module."):]: v for k, v in full.items() if k.startswith("torch_nn_module.")},
                          strict=False)
    z_f = LatentDiffusion(torch_nn_module=fresh).ddim_sample_loop(cond=cond, shape=shape, x_T=zT, ddim_steps=10)
    assert torch.equal(z_c, z_f)
    # eager (no graph) == graph
    os.environ["PD_NO_GRAPH"] = "1"
    try:
        z_e = ldm.ddim_sample_loop(cond=cond, shape=shape, x_T=zT, ddim_steps=10)
    finally:
        del os.environ["PD_NO_GRAPH"]
    assert torch.equal(z_e, z_c)


def test_ema_scope_swaps_the_weights_the_device_loop_uses():
    unet, _ = make_unet(CFG)
    ldm = LatentDiffusion(torch_nn_module=unet, use_ema=True)
    # drift the shadow weights away from the live ones
    g = torch.Generator().manual_seed(5)
    for name, buf in ldm.model_ema.named_buffers():
        if buf.dtype == torch.float32 and buf.dim() >= 2:
            buf.mul_(1.0 + 0.05 * torch.randn(buf.shape, generator=g))
    zT, cond, _ = _tiny_inputs()
    shape = tuple(zT.shape)
    z_live = ldm.ddim_sample_loop(cond=cond, shape=shape, x_T=zT, ddim_steps=10)
    with ldm.ema_scope():
        z_ema = ldm.ddim_sample_loop(cond=cond, shape=shape, x_T=zT, ddim_steps=10)
        ema_sd = {k: v.detach().clone() for k, v in unet.state_dict().items()}
    z_back = ldm.ddim_sample_loop(cond=cond, shape=shape, x_T=zT, ddim_steps=10)
    assert torch.equal(z_back, z_live) and not torch.equal(z_ema, z_live)
    fresh, _ = make_unet(CFG)
    fresh.load_state_dict(ema_sd, strict=False)
    z_f = LatentDiffusion(torch_nn_module=fresh).ddim_sample_loop(cond=cond, shape=shape, x_T=zT, ddim_steps=10)
    assert torch.equal(z_ema, z_f)


def test_full_size_sample_vs_reference():
    """Shipped sizes end to end, batch 1: 7 frames 128x128 -> AutoencoderKL.encode -> loop -> decode -> 6 frames.
    (a) the reference's own sample() with 4 ancestral steps (RNG injected); (b) sample(sampler='ddim'): the same encode /
    decode around the S6 50-step DDIM, golden = reference modules + the reference's DDIM helpers."""
    fu, fv = Wt.UNetConfig(), Wt.VAEConfig()
    unet, _ = make_unet(fu, max_batch=1)
    vae, _ = make_vae(fv, max_frames=8)
    ldm = LatentDiffusion(torch_nn_module=unet, first_stage_model=vae, cond_stage_model="__is_first_stage__")
    y = inp(880, 1, fu.t_in, fv.h, fv.w, 1, uniform=True).cuda()
    zT = inp(881, 1, fu.t_out, fu.h, fu.w, fu.c).cuda()
    noise = inp(882, 4, 1, fu.t_out, fu.h, fu.w, fu.c).cuda()
    zc = ldm.cond_stage_forward({"y": y})
    rz, mz = errs(zc, G["full_zc"])
    dec = ldm.sample(cond={"y": y}, batch_size=1, x_T=zT, timesteps=4, noise=noise)
    assert tuple(dec.shape) == (1, 6, 128, 128, 1)
    r4, m4 = errs(dec, G["full_sample_ddpm4"])
    dec50 = ldm.sample(cond={"y": y}, batch_size=1, x_T=zT, sampler="ddim", ddim_steps=50)
    r50, m50 = errs(dec50, G["full_sample_ddim50"])
    print(f"full-size sample(): zc rel_rms={rz:.3e}; ddpm-4 frames rel_rms={r4:.3e} max={m4:.3e}; "
          f"ddim-50 frames rel_rms={r50:.3e} max={m50:.3e}")
    assert rz < 1.5e-2 and r4 < LOOP_RMS_TOL and m4 < LOOP_MAX_TOL and r50 < FRAME_RMS_TOL and m50 < FRAME_MAX_TOL
