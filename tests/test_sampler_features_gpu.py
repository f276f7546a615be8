"""p_sample_loop call-surface features of the reference (latent_diffusion.py:633-684) on the CUDA path - SURVEY.md
section 8f rank 1: the reference's per-step torch.randn stream, `x_T=None`, `start_T`, `callback` / `img_callback`,
`mask` / `x0` inpainting (host-driven), and the bit-exact invariances DESIGN.md section 2 claims (shard invariance,
graph replay == eager, re-run determinism)."""
import os

import pytest
import torch

from prediff_b200 import weights as Wt
from prediff_b200.diffusion import LatentDiffusion
from tests.golden.gen_golden import inp
from tests.test_unet_gpu import make_unet

pytestmark = pytest.mark.gpu
CFG = Wt.TINY_UNET


@pytest.fixture(scope="module")
def ldm():
    unet, _ = make_unet(CFG, max_batch=4)
    return LatentDiffusion(torch_nn_module=unet, data_shape=(6, 128, 128, 1), latent_shape=(6, 16, 16, 64),
                           first_stage_model=None, cond_stage_model=None)


def _cond(B, seed=778):
    return inp(seed, B, CFG.t_in, CFG.h, CFG.w, CFG.c).cuda()


def test_default_noise_is_the_reference_randn_stream(ldm):
    """x_T = randn(shape) first, then one randn(shape) per step in loop order (latent_diffusion.py:645-648, 620)."""
    shape = (2, CFG.t_out, CFG.h, CFG.w, CFG.c)
    cond = _cond(2)
    torch.manual_seed(4321)
    a = ldm.p_sample_loop(cond=cond, shape=shape, timesteps=5)
    torch.manual_seed(4321)
    x_T = torch.randn(shape, device="cuda")
    noise = torch.stack([torch.randn(shape, device="cuda") for _ in range(5)])
    b = ldm.p_sample_loop(cond=cond, shape=shape, x_T=x_T, timesteps=5, noise=noise)
    assert torch.equal(a, b)
    # the host-driven route (what alignment / inpainting use) draws the same stream
    torch.manual_seed(4321)
    x_T2 = torch.randn(shape, device="cuda")
    img = x_T2
    for i in reversed(range(5)):
        img = ldm.p_sample(zt=img, zc=cond, t=torch.full((2,), i, device="cuda"))
    assert torch.equal(x_T2, x_T)
    # same noise, same kernels: p_sample runs through pd_sample_step_ddpm (UNet + the fused update), so the host-driven
    # route reproduces the device-resident loop bit for bit
    assert torch.equal(img, a)


def test_start_T_callbacks_and_intermediates(ldm):
    shape = (2, CFG.t_out, CFG.h, CFG.w, CFG.c)
    cond, x_T = _cond(2), inp(777, *shape).cuda()
    noise = inp(779, 6, *shape).cuda()
    seen, imgs = [], []
    z, inter = ldm.p_sample_loop(cond=cond, shape=shape, x_T=x_T, timesteps=1000, start_T=6, noise=noise, log_every_t=2,
                                 return_intermediates=True, callback=seen.append,
                                 img_callback=lambda im, i: imgs.append((i, im.clone())))
    assert seen == [5, 4, 3, 2, 1, 0]                       # start_T truncates the chain to t = 5..0
    assert [i for i, _ in imgs] == seen and torch.equal(imgs[-1][1], z)
    # intermediates: z_T, then every step with i % log_every_t == 0 or i == timesteps - 1 (:677-678) -> t = 5, 4, 2, 0
    assert len(inter) == 5 and torch.equal(inter[0], x_T) and torch.equal(inter[-1], z)
    assert torch.equal(inter[1], imgs[0][1]) and torch.equal(inter[2], imgs[1][1]) and torch.equal(inter[3], imgs[3][1])
    # no callbacks: one device-resident stretch gives the same bits
    z2 = ldm.p_sample_loop(cond=cond, shape=shape, x_T=x_T, timesteps=6, noise=noise)
    assert torch.equal(z2, z)


def test_inpainting_mask_keeps_known_region(ldm):
    """mask / x0 (latent_diffusion.py:673-675): img = q_sample(x0, t) * mask + (1 - mask) * img after every step."""
    shape = (2, CFG.t_out, CFG.h, CFG.w, CFG.c)
    cond, x_T = _cond(2), inp(777, *shape).cuda()
    x0 = inp(555, *shape).cuda()
    mask = torch.zeros(shape, device="cuda")
    mask[:, :, :8] = 1.0
    torch.manual_seed(7)
    z = ldm.p_sample_loop(cond=cond, shape=shape, x_T=x_T, timesteps=4, mask=mask, x0=x0)
    ac0 = ldm.alphas_cumprod[0].item()
    known = z[:, :, :8]
    # at t = 0 the known region is sqrt(ac_0) x0 + sqrt(1 - ac_0) n: within 6 sigma of the scaled x0 everywhere
    assert (known - ac0 ** 0.5 * x0[:, :, :8]).abs().max().item() < 6 * (1 - ac0) ** 0.5
    assert (z[:, :, 8:] - ac0 ** 0.5 * x0[:, :, 8:]).abs().max().item() > 0.5   # the free region is generated
    assert torch.isfinite(z).all()


def test_shard_invariance_graph_vs_eager_and_determinism(ldm):
    """Rows [2,4) sampled alone equal the same rows inside the batch of 4; CUDA-graph replay == eager launches."""
    shape4 = (4, CFG.t_out, CFG.h, CFG.w, CFG.c)
    cond, x_T = _cond(4), inp(777, *shape4).cuda()
    full = ldm.ddim_sample_loop(cond=cond, shape=shape4, x_T=x_T, ddim_steps=5)
    part = ldm.ddim_sample_loop(cond=cond[2:].contiguous(), shape=(2,) + shape4[1:], x_T=x_T[2:].contiguous(), ddim_steps=5)
    assert torch.equal(full[2:], part)
    again = ldm.ddim_sample_loop(cond=cond, shape=shape4, x_T=x_T, ddim_steps=5)
    assert torch.equal(again, full)
    os.environ["PD_NO_GRAPH"] = "1"
    try:
        eager = ldm.ddim_sample_loop(cond=cond, shape=shape4, x_T=x_T, ddim_steps=5)
    finally:
        del os.environ["PD_NO_GRAPH"]
    assert torch.equal(eager, full)
    for nsub in ("2", "4"):   # the sub-batch split is an execution detail: same bits
        os.environ["PD_SUB_BATCHES"] = nsub
        try:
            other = ldm.ddim_sample_loop(cond=cond, shape=shape4, x_T=x_T, ddim_steps=5)
        finally:
            del os.environ["PD_SUB_BATCHES"]
        assert torch.equal(other, full)
