"""On-device evaluation step: SEVIRSkillScore contingency counts + MSE / MAE sums.

CPU: the numpy oracle (oracle/eval_oracle.py) against goldens of the unmodified reference's SEVIRSkillScore
(tests/golden/skill.npz). GPU: the CUDA kernel through the C ABI / Python mirror against the goldens and the oracle -
integer counts bit-exact, scores to fp32 round-off."""
import os

import numpy as np
import pytest
import torch

from oracle import eval_oracle as EO
from tests.golden.gen_golden import skill_inputs

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "skill.npz"))
THR = (16, 74, 133, 160, 181, 219)
METRICS = ("csi", "bias", "sucr", "pod")


def _golden_counts(tag, mode):
    return np.stack([G[f"{tag}_m{mode}_{k}"] for k in ("hits", "misses", "fas")], axis=-1).astype(np.int64)


@pytest.mark.parametrize("tag,pool", [("p1", 1), ("p4", 4)])
def test_oracle_counts_and_scores_match_reference(tag, pool):
    pred, target = skill_inputs()
    c = EO.hits_misses_fas(pred.numpy(), target.numpy(), THR, pool) + \
        EO.hits_misses_fas(pred.flip(0).numpy(), target.numpy(), THR, pool)
    assert np.array_equal(c, _golden_counts(tag, "1"))                 # (n_thr, T, 3), per lead time
    assert np.array_equal(c.sum(axis=1), _golden_counts(tag, "0"))     # mode "0": summed over T
    for mode in ("0", "1", "2"):
        res = EO.scores(c, THR, METRICS, mode)
        for m in METRICS:
            got = np.stack([np.asarray(res[t][m], dtype=np.float64) for t in THR])
            np.testing.assert_allclose(got, G[f"{tag}_m{mode}_{m}"], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(np.asarray(res["avg"][m], dtype=np.float64), G[f"{tag}_m{mode}_avg_{m}"], rtol=2e-6,
                                       atol=1e-7)


def test_oracle_edge_cases():
    z = np.zeros((1, 2, 4, 4), dtype=np.float32)
    assert EO.hits_misses_fas(z, z, THR).sum() == 0                     # nothing above any threshold
    one = np.ones((1, 2, 4, 4), dtype=np.float32)
    c = EO.hits_misses_fas(one, one, THR)
    assert (c[..., 0] == 16).all() and c[..., 1:].sum() == 0            # all hits
    exact = np.full((1, 1, 4, 4), np.float32(74) / np.float32(255), dtype=np.float32)
    c = EO.hits_misses_fas(exact, z[:, :1], (74,))
    assert c[0, 0].tolist() == [0, 0, int((EO.back_transform(exact) >= 74).sum())]   # >= at the exact threshold
    assert EO.hits_misses_fas(np.zeros((0, 2, 4, 4), np.float32), np.zeros((0, 2, 4, 4), np.float32), THR).sum() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("tag,pre", [("p1", "sevir"), ("p4", "sevir_pool4")])
def test_gpu_skill_score_matches_reference_golden(tag, pre):
    from prediff_b200.evaluation import SEVIRSkillScore
    pred, target = skill_inputs()
    for mode in ("0", "1", "2"):
        sc = SEVIRSkillScore(layout="NTHWC", mode=mode, seq_len=6, preprocess_type=pre, threshold_list=THR,
                             metrics_list=METRICS, eps=1e-4)
        sc.update(pred.unsqueeze(-1).cuda(), target.unsqueeze(-1).cuda())
        sc.update(pred.flip(0).unsqueeze(-1).cuda(), target.unsqueeze(-1).cuda())
        assert np.array_equal(sc.hits_misses_fas, _golden_counts(tag, "1"))   # bit-exact integer state
        res = sc.compute()
        for m in METRICS:
            got = np.stack([np.asarray(res[t][m], dtype=np.float64) for t in THR])
            np.testing.assert_allclose(got, G[f"{tag}_m{mode}_{m}"], rtol=2e-6, atol=1e-7)
            np.testing.assert_allclose(np.asarray(res["avg"][m], dtype=np.float64), G[f"{tag}_m{mode}_avg_{m}"], rtol=2e-6,
                                       atol=1e-7)
        sc.reset()
        sc.update(torch.nan_to_num(pred).unsqueeze(-1).cuda(), torch.nan_to_num(target).unsqueeze(-1).cuda())
        assert abs(sc.mse() - float(G["mse_nonan"])) < 1e-9 + 1e-6 * float(G["mse_nonan"])
        assert abs(sc.mae() - float(G["mae_nonan"])) < 1e-9 + 1e-6 * float(G["mae_nonan"])


@pytest.mark.gpu
@pytest.mark.parametrize("N,T,H,W,pool,layout", [(4, 6, 128, 128, 1, "NTHWC"), (2, 6, 128, 128, 16, "NTHW"),
                                                (3, 5, 48, 80, 4, "NHWT"), (1, 1, 8, 8, 1, "NTHW"),
                                                (0, 6, 16, 16, 1, "NTHW")])
def test_gpu_counts_match_oracle_random(N, T, H, W, pool, layout):
    """Full-size frames (the decoder's 128 x 128 output), pooled variants, other layouts, empty batch: bit-exact
    against the oracle; the checksum property hits + misses == #(target >= thr) holds by construction."""
    from prediff_b200.evaluation import SEVIRSkillScore
    rng = np.random.Generator(np.random.PCG64(99 + N + H))
    target = rng.random((N, T, H, W), dtype=np.float32)
    pred = np.clip(target + 0.2 * rng.standard_normal((N, T, H, W), dtype=np.float32), 0, 1).astype(np.float32)
    if N:
        pred[0, 0, 0, 0] = np.nan
    pre = "sevir" if pool == 1 else f"sevir_pool{pool}"
    sc = SEVIRSkillScore(layout=layout, mode="1", seq_len=T, preprocess_type=pre, threshold_list=THR)

    def to_layout(a):
        t = torch.from_numpy(a)
        if "C" in layout:
            t = t.unsqueeze(-1)
        src = "NTHWC" if "C" in layout else "NTHW"
        return t.permute(*[src.find(ax) for ax in layout]).contiguous().cuda()

    if N == 0:
        sc.update(to_layout(pred), to_layout(target))
        assert sc.hits_misses_fas.sum() == 0
        return
    sc.update(to_layout(pred), to_layout(target))
    ref = EO.hits_misses_fas(pred, target, THR, pool)
    assert np.array_equal(sc.hits_misses_fas, ref)
    tb = EO.max_pool(EO.back_transform(target), pool)
    pb = EO.max_pool(EO.back_transform(pred), pool)
    ok = ~(np.isnan(tb) | np.isnan(pb))
    for i, th in enumerate(THR):
        assert (ref[i, :, 0] + ref[i, :, 1]).sum() == ((tb >= th) & ok).sum()


# ---- SSIM (torchmetrics 1.2.0 StructuralSimilarityIndexMeasure defaults; restated algorithm, see oracle/eval_oracle.py) ----
def _ssim_direct(p, t, data_range):
    """The definition evaluated with plain double loops (float64): Gaussian-weighted local moments over the 11 x 11 window,
    interior pixels only - an independent check of the oracle's conv2d / padding / cropping arithmetic."""
    g = np.exp(-((np.arange(11) - 5) / 1.5) ** 2 / 2)
    g = g / g.sum()
    w = np.outer(g, g)
    H, W = p.shape
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    vals = []
    for y in range(5, H - 5):
        for x in range(5, W - 5):
            a, b = p[y - 5:y + 6, x - 5:x + 6], t[y - 5:y + 6, x - 5:x + 6]
            ma, mb = (w * a).sum(), (w * b).sum()
            va, vb, vab = (w * a * a).sum() - ma * ma, (w * b * b).sum() - mb * mb, (w * a * b).sum() - ma * mb
            vals.append(((2 * ma * mb + c1) * (2 * vab + c2)) / ((ma * ma + mb * mb + c1) * (va + vb + c2)))
    return float(np.mean(vals))


def _ssim_inputs(seed=808, B=3, H=40, W=36):
    rng = np.random.Generator(np.random.PCG64(seed))
    t = rng.random((B, 1, H, W), dtype=np.float32)
    p = np.clip(t + 0.2 * rng.standard_normal((B, 1, H, W)).astype(np.float32), 0, 1).astype(np.float32)
    return torch.from_numpy(p), torch.from_numpy(t)


def test_ssim_oracle_known_answers():
    p, t = _ssim_inputs()
    assert torch.allclose(EO.ssim_per_image(t, t), torch.ones(3), atol=1e-6)          # identical images
    s = EO.ssim_per_image(p, t)
    dr = float(max(p.max() - p.min(), t.max() - t.min()))
    for b in range(3):
        assert abs(float(s[b]) - _ssim_direct(p[b, 0].double().numpy(), t[b, 0].double().numpy(), dr)) < 2e-5
    # the reflect-padded border is cropped away again: changing what the padding mode would see changes nothing as long
    # as the pixels themselves are the same (the interior only reads real pixels)
    s_fixed = EO.ssim_per_image(p, t, data_range=1.0)
    assert abs(float(s_fixed[0]) - _ssim_direct(p[0, 0].double().numpy(), t[0, 0].double().numpy(), 1.0)) < 2e-5
    st = EO.SSIMState()
    st.update(p[:2], t[:2])
    st.update(p[2:], t[2:])   # data_range is taken per update batch
    want = (float(EO.ssim_per_image(p[:2], t[:2]).sum()) + float(EO.ssim_per_image(p[2:], t[2:]).sum())) / 3
    assert abs(st.compute() - want) < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 40, 36), (12, 128, 128), (1, 11 + 32, 11 + 33)])
def test_gpu_ssim_matches_oracle(shape):
    from prediff_b200.evaluation import StructuralSimilarityIndexMeasure
    B, H, W = shape
    p, t = _ssim_inputs(seed=909, B=B, H=H, W=W)
    m = StructuralSimilarityIndexMeasure()
    o = EO.SSIMState()
    for lo, hi in ((0, max(1, B // 2)), (max(1, B // 2), B)):
        if hi > lo:
            m(p[lo:hi].cuda(), t[lo:hi].cuda())
            o.update(p[lo:hi], t[lo:hi])
    got, want = float(m.compute()), o.compute()
    print(f"ssim {shape}: cuda {got:.7f} oracle {want:.7f}")
    assert abs(got - want) < 2e-5
    fixed = StructuralSimilarityIndexMeasure(data_range=1.0)
    fixed.update(p.cuda(), t.cuda())
    assert abs(float(fixed.compute()) - float(EO.ssim_per_image(p, t, data_range=1.0).mean())) < 2e-5
    same = StructuralSimilarityIndexMeasure()
    same.update(t.cuda(), t.cuda())
    assert abs(float(same.compute()) - 1.0) < 1e-6
    with pytest.raises(NotImplementedError):
        StructuralSimilarityIndexMeasure(kernel_size=7)
