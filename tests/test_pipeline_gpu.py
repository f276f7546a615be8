"""The evaluation pass of the inference script as one device-resident pipeline (SURVEY.md 8f ranks 2 + 3 chained around the
sampling path; train_sevirlr_prediff.py:924-965 `test_step`): raw uint8 events -> SEVIRDataLoader windows (pinned staging,
copy stream, window kernel) -> context / target split -> LatentDiffusion.sample (encode, device-resident DDIM loop, decode) ->
SEVIRSkillScore / MSE / MAE / SSIM accumulated by the evaluation kernels. Nothing crosses to the host between the uint8 events
and the final scores; the checks below read tensors back only to compare them with the oracles:
  - the metric states equal the CPU oracles (oracle/eval_oracle.py) evaluated on the frames the pipeline produced (integer
    counts bit-exact, SSIM / MSE to fp32 round-off);
  - the forecasts equal a stand-alone sample() on the same context (the loader's recycled buffers and streams do not race
    the sampler), and the whole pass is deterministic."""
import numpy as np
import pytest
import torch

from oracle import data_oracle as DO
from oracle import eval_oracle as EO
from prediff_b200 import weights as Wt
from prediff_b200.data import SEVIRDataLoader
from prediff_b200.diffusion import LatentDiffusion
from prediff_b200.evaluation import SEVIRSkillScore, StructuralSimilarityIndexMeasure
from tests.test_unet_gpu import make_unet
from tests.test_vae_gpu import make_vae

pytestmark = pytest.mark.gpu
THR = (16, 74, 133, 160, 181, 219)


def run_pass(ldm, ev, batch_size, keep):
    dl = SEVIRDataLoader(ev, seq_len=13, stride=6, batch_size=batch_size, layout="NTHWC", rescale_method="01", prefetch=2)
    skill = SEVIRSkillScore(layout="NTHWC", mode="1", seq_len=6, preprocess_type="sevir", threshold_list=THR)
    ssim = StructuralSimilarityIndexMeasure()
    g = torch.Generator(device="cuda").manual_seed(123)
    for i, seq in enumerate(dl):                       # (B, 13, 128, 128, 1) fp32 on the device, valid until the next step
        ctx, tgt = seq[:, :7], seq[:, 7:]
        x_T = torch.randn(batch_size, 6, 16, 16, 64, device="cuda", generator=g)
        pred = ldm.sample(cond={"y": ctx.contiguous()}, batch_size=batch_size, x_T=x_T, sampler="ddim", ddim_steps=4)
        skill.update(pred, tgt)
        ssim.update(pred.permute(0, 1, 4, 2, 3).reshape(-1, 1, 128, 128), tgt.permute(0, 1, 4, 2, 3).reshape(-1, 1, 128, 128))
        keep.append((pred.clone(), tgt.clone(), ctx.clone(), x_T))
    return skill, ssim, len(dl)


def test_loader_sample_metrics_pipeline_stays_on_device_and_matches_oracles():
    unet, _ = make_unet(Wt.TINY_UNET)
    vae, _ = make_vae(Wt.TINY_VAE)
    ldm = LatentDiffusion(torch_nn_module=unet, data_shape=(6, 128, 128, 1), latent_shape=(6, 16, 16, 64),
                          first_stage_model=vae, cond_stage_model="__is_first_stage__")
    rng = np.random.Generator(np.random.PCG64(2024))
    # smooth synthetic VIL events: a blob drifting over a noisy background, uint8 'NHWT'
    yy, xx = np.mgrid[0:128, 0:128].astype(np.float32)
    ev = np.empty((3, 128, 128, 25), np.uint8)
    for e in range(3):
        for t in range(25):
            cx, cy = 30 + 3 * t + 10 * e, 90 - 2 * t
            blob = 220 * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * 18.0 ** 2))
            ev[e, :, :, t] = np.clip(blob + 12 * rng.random((128, 128), dtype=np.float32), 0, 255).astype(np.uint8)
    keep = []
    skill, ssim, n = run_pass(ldm, ev, 2, keep)
    assert n == (3 * 3) // 2 == len(keep)
    # windows: bit-exact against the reference rule
    for i, (pred, tgt, ctx, _) in enumerate(keep):
        want = DO.idx_sample(ev, i, 2, 13, 6, "01", "NTHWC")
        assert np.array_equal(torch.cat([ctx, tgt], 1).cpu().numpy(), want)
        assert tuple(pred.shape) == (2, 6, 128, 128, 1) and torch.isfinite(pred).all()
    # metric states == oracles on the very frames the pipeline produced
    counts = sum(EO.hits_misses_fas(p.cpu().numpy()[..., 0], t.cpu().numpy()[..., 0], THR) for p, t, _, _ in keep)
    assert np.array_equal(skill.hits_misses_fas, counts)
    se = sum(float(((p.double() - t.double()) ** 2).sum()) for p, t, _, _ in keep)
    numel = sum(p.numel() for p, _, _, _ in keep)
    assert abs(skill.mse() - se / numel) < 1e-5 * se / numel
    o = EO.SSIMState()
    for p, t, _, _ in keep:
        o.update(p.permute(0, 1, 4, 2, 3).reshape(-1, 1, 128, 128).cpu(), t.permute(0, 1, 4, 2, 3).reshape(-1, 1, 128, 128).cpu())
    assert abs(float(ssim.compute()) - o.compute()) < 2e-5
    sc = skill.compute()
    assert set(sc) == set(THR) | {"avg"} and sc["avg"]["csi"].shape == (6,)
    # the forecasts equal a stand-alone sample() on the same context / z_T; a second pass repeats the first bit for bit
    for pred, _, ctx, x_T in keep:
        again = ldm.sample(cond={"y": ctx}, batch_size=2, x_T=x_T, sampler="ddim", ddim_steps=4)
        assert torch.equal(again, pred)
    keep2 = []
    skill2, ssim2, _ = run_pass(ldm, ev, 2, keep2)
    assert np.array_equal(skill2.hits_misses_fas, skill.hits_misses_fas)
    assert float(ssim2.compute()) == float(ssim.compute())
