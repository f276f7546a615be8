"""Generates the golden fixtures in this directory by running the UNMODIFIED reference modules on CPU.

Run in the build container only (needs /root/reference; the GPU box does not have it):
    python tests/golden/gen_golden.py [--only schedule,unet_tiny,...] [--ddim-full]

`lightning` and `diffusers` are absent from the image; they contribute no arithmetic to the sampling path
(SURVEY.md section 8c), so they are replaced by inert sys.modules stubs before importing the reference.
Weights come from prediff_b200.weights.seeded_state_dict (portable numpy PCG64 streams) and are loaded into the
reference modules with load_state_dict; inputs come from the same kind of streams (`inp()` below), so tests can
regenerate them bit-exactly anywhere. Only outputs (and small inputs) are stored.
"""
import argparse
import datetime
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("PREDIFF_REFERENCE", "/root/reference")

from prediff_b200 import weights as Wt  # noqa: E402

UNET_SEED, VAE_SEED, KA_SEED = 1001, 2002, 3003


def inp(seed, *shape, uniform=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.random(shape, dtype=np.float32) if uniform else rng.standard_normal(shape, dtype=np.float32)
    return torch.from_numpy(a)


def install_stubs():
    import torch.nn as nn

    class _LM(nn.Module):
        @property
        def device(self):
            return next(self.parameters()).device if any(True for _ in self.parameters()) else torch.device("cpu")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("lightning")
    mod("lightning.pytorch", LightningModule=_LM)
    mod("lightning.pytorch.utilities")
    mod("lightning.pytorch.utilities.rank_zero", rank_zero_only=lambda f: f)
    sys.modules["lightning"].pytorch = sys.modules["lightning.pytorch"]
    mod("diffusers")
    mod("diffusers.models")
    mod("diffusers.models.autoencoder_kl", AutoencoderKLOutput=type("AutoencoderKLOutput", (), {}),
        DecoderOutput=type("DecoderOutput", (), {}))
    sys.path.insert(0, os.path.join(REF, "src"))


def ref_unet(cfg):
    from prediff.models.cuboid_transformer import CuboidTransformerUNet
    pats = list(getattr(cfg, "patterns", ("axial", "axial")))
    explicit = {}
    if getattr(cfg, "explicit_layers", None) is not None:   # block_attn_patterns=None + explicit per-block lists
        pats = None
        explicit = dict(block_cuboid_size=[[a for a, _, _ in blk] for blk in cfg.explicit_layers],
                        block_cuboid_strategy=[[b for _, b, _ in blk] for blk in cfg.explicit_layers],
                        block_cuboid_shift_size=[[c for _, _, c in blk] for blk in cfg.explicit_layers])
    # arguments as in scripts/prediff/sevirlr/train_sevirlr_prediff.py:91-137 with cfg.yaml:157-206
    m = CuboidTransformerUNet(
        input_shape=[cfg.t_in, cfg.h, cfg.w, cfg.c], target_shape=[cfg.t_out, cfg.h, cfg.w, cfg.c],
        base_units=cfg.base_units, scale_alpha=1.0, num_heads=cfg.num_heads, attn_drop=0.1, proj_drop=0.1, ffn_drop=0.1,
        downsample=2, downsample_type="patch_merge", upsample_type="upsample", upsample_kernel_size=3,
        depth=list(cfg.depth), block_attn_patterns=pats, num_global_vectors=getattr(cfg, "num_global_vectors", 0),
        use_global_vector_ffn=getattr(cfg, "use_global_vector_ffn", True) if getattr(cfg, "num_global_vectors", 0) else False,
        use_global_self_attn=getattr(cfg, "use_global_self_attn", False) if getattr(cfg, "num_global_vectors", 0) else True,
        separate_global_qkv=getattr(cfg, "separate_global_qkv", False) if getattr(cfg, "num_global_vectors", 0) else True,
        global_dim_ratio=1, ffn_activation="gelu", gated_ffn=False,
        norm_layer="layer_norm", padding_type=getattr(cfg, "padding_type", "zeros"), checkpoint_level=0,
        pos_embed_type="t+h+w", use_relative_pos=True, self_attn_use_final_proj=True, time_embed_channels_mult=4,
        time_embed_use_scale_shift_norm=False, time_embed_dropout=0.0, unet_res_connect=True, **explicit)
    sd = Wt.seeded_state_dict(Wt.unet_param_spec(cfg), UNET_SEED)
    res = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.endswith("relative_position_index") for k in res.missing_keys), res.missing_keys
    ref_keys = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.endswith("relative_position_index")}
    assert ref_keys == {k: tuple(v.shape) for k, v in sd.items()}, "spec != reference state_dict"
    assert list(ref_keys) == [k for k, _ in Wt.unet_param_spec(cfg)], "spec order != reference registration order"
    # the derived buffer must equal the reference's
    for k, v in m.state_dict().items():
        if k.endswith("relative_position_index"):
            lvl = int(k.split(".")[1])
            i = int(k.split(".")[4])
            assert np.array_equal(v.numpy(), Wt.relative_position_index(cfg.cuboids(lvl)[i])), k
    return m.eval()


def ref_vae(cfg):
    from prediff.taming import AutoencoderKL
    m = AutoencoderKL(down_block_types=["DownEncoderBlock2D"] * 4, in_channels=cfg.in_channels,
                      block_out_channels=list(cfg.block_out_channels), act_fn="silu",
                      latent_channels=cfg.latent_channels, up_block_types=["UpDecoderBlock2D"] * 4,
                      norm_num_groups=cfg.norm_num_groups, layers_per_block=cfg.layers_per_block,
                      out_channels=cfg.out_channels)
    sd = Wt.seeded_state_dict(Wt.vae_param_spec(cfg), VAE_SEED)
    ref_keys = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert ref_keys == {k: tuple(v.shape) for k, v in sd.items()}, \
        (set(ref_keys) ^ set(sd), [k for k in ref_keys if k in sd and ref_keys[k] != sd[k].shape])
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    return m.eval()


def ref_ldm(unet, vae, ucfg, vcfg, **kw):
    from prediff.diffusion.latent_diffusion import LatentDiffusion
    kw.setdefault("learn_logvar", False)
    return LatentDiffusion(
        **kw,
        torch_nn_module=unet, layout="NTHWC", data_shape=(ucfg.t_out, vcfg.h, vcfg.w, 1), timesteps=1000,
        beta_schedule="linear", use_ema=False, log_every_t=100, clip_denoised=False, linear_start=1e-4,
        linear_end=2e-2, parameterization="eps",
        latent_shape=(ucfg.t_out, ucfg.h, ucfg.w, ucfg.c), first_stage_model=vae,
        cond_stage_model="__is_first_stage__", scale_factor=1.0).eval()


def ref_ka(cfg):
    from prediff.diffusion.knowledge_alignment.sevir import SEVIRAvgIntensityAlignment
    # model_args as in scripts/prediff/sevirlr/cfg.yaml:105-156
    args = dict(input_shape=[cfg.t, cfg.h, cfg.w, cfg.c], out_channels=1, base_units=cfg.base_units, scale_alpha=1.0,
                depth=list(cfg.depth), downsample=2, downsample_type="patch_merge", block_attn_patterns="axial",
                num_heads=cfg.num_heads, attn_drop=0.1, proj_drop=0.1, ffn_drop=0.1, ffn_activation="gelu", gated_ffn=False,
                norm_layer="layer_norm", use_inter_ffn=True, hierarchical_pos_embed=False, pos_embed_type="t+h+w",
                padding_type="zeros", checkpoint_level=0, use_relative_pos=True, self_attn_use_final_proj=True,
                num_global_vectors=0, use_global_vector_ffn=True, use_global_self_attn=False, separate_global_qkv=False,
                global_dim_ratio=1, time_embed_channels_mult=4, time_embed_use_scale_shift_norm=False,
                time_embed_dropout=0.0, pool="attention", readout_seq=True, out_len=cfg.t)
    al = SEVIRAvgIntensityAlignment(alignment_type="avg_x", guide_scale=cfg.guide_scale, model_type="cuboid",
                                    model_args=args, model_ckpt_path=None)
    sd = Wt.seeded_state_dict(Wt.ka_param_spec(cfg), KA_SEED)
    res = al.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.endswith("relative_position_index") for k in res.missing_keys), res.missing_keys
    ref_keys = {k: tuple(v.shape) for k, v in al.model.state_dict().items() if not k.endswith("relative_position_index")}
    assert ref_keys == {k: tuple(v.shape) for k, v in sd.items()}, set(ref_keys) ^ set(sd)
    assert list(ref_keys) == [k for k, _ in Wt.ka_param_spec(cfg)]
    al.model.eval()
    return al


def gen_ka():
    """U(z_t, t) and the guidance g = guide_scale * grad || mean_T U - avg_x_gt ||_2 of the reference
    (SEVIRAvgIntensityAlignment.get_mean_shift), plus one aligned DDPM step of the reference LatentDiffusion."""
    import prediff.diffusion.latent_diffusion as LD
    cfg = Wt.KAConfig()
    al = ref_ka(cfg)
    B = 4
    zt = inp(5151, B, cfg.t, cfg.h, cfg.w, cfg.c)
    t = torch.tensor([981, 500, 20, 0], dtype=torch.long)
    target = torch.full((B, 1), 0.3)
    with torch.no_grad():
        pred = al.model(zt, t)
    g = al.get_mean_shift(zt, t, avg_x_gt=target)
    print(f"ka: pred std {pred.std():.4f}, grad absmax {g.abs().max():.4e}")
    # aligned p_sample of the reference LatentDiffusion (tiny UNet) with the full KA network
    ucfg = Wt.TINY_UNET
    ldm = ref_ldm(ref_unet(ucfg), ref_vae(Wt.TINY_VAE), ucfg, Wt.TINY_VAE)
    ldm.set_alignment(al.get_mean_shift)
    zT = inp(777, 2, ucfg.t_out, ucfg.h, ucfg.w, ucfg.c)
    cond = inp(778, 2, ucfg.t_in, ucfg.h, ucfg.w, ucfg.c)
    noise = inp(779, 4, 2, ucfg.t_out, ucfg.h, ucfg.w, ucfg.c)
    orig = LD.noise_like
    LD.noise_like = lambda shape, device: noise[0].clone()
    try:
        ts = torch.full((2,), 900, dtype=torch.long)
        z_al = ldm.p_sample(zt=zT.clone(), zc=cond, t=ts, use_alignment=True,
                            alignment_kwargs={"avg_x_gt": torch.full((2, 1), 0.3)})
    finally:
        LD.noise_like = orig
    save("ka_full", t=t, pred=pred, grad=g, z_aligned_step900=z_al)


def skill_inputs(seed=6161, N=3, T=6, H=32, W=32):
    """Smooth-ish synthetic forecast / target pairs in [0,1] with values on both sides of every threshold, a few
    exact-threshold pixels (k/255) and a few NaNs."""
    rng = np.random.Generator(np.random.PCG64(seed))
    target = rng.random((N, T, H, W), dtype=np.float32) ** 2
    pred = np.clip(target + 0.15 * rng.standard_normal((N, T, H, W), dtype=np.float32), 0.0, 1.0).astype(np.float32)
    for i, th in enumerate((16, 74, 133, 160, 181, 219)):
        target[0, i % T, i, :8] = np.float32(th) / np.float32(255.0)
        pred[1, i % T, :8, i] = np.float32(th) / np.float32(255.0)
    pred[2, 1, 3, 4] = np.nan
    target[2, 2, 5, 6] = np.nan
    target[0, 0, 30, 30] = np.nan
    pred[0, 0, 30, 30] = np.nan
    return torch.from_numpy(pred), torch.from_numpy(target)


def loader_events(seed=7171, E=5, H=16, W=24, T_raw=25):
    """Synthetic raw VIL events, uint8 'NHWT' like the SEVIR-LR HDF5 datasets (sevir_dataloader.py:22, :360-380)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 256, size=(E, H, W, T_raw), dtype=np.uint8)


def gen_loader():
    """Batches of the unmodified reference SEVIRDataLoader._idx_sample (sequent windows + preprocess + layout change,
    sevir_dataloader.py:834-891, :610-650). The HDF5 read (_load_event_batch) is replaced by a slice of a synthetic
    uint8 event array - h5py / the dataset are absent - everything after it is the reference's own code."""
    for name in ("h5py",):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    from prediff.datasets.sevir.sevir_dataloader import SEVIRDataLoader
    ev = loader_events()
    out = {}
    for tag, (bs, seq_len, stride, rescale, layout) in {
        "lr": (4, 13, 6, "01", "NTHWC"),          # shipped SEVIR-LR config (cfg.yaml:5-12)
        "b3": (3, 13, 6, "01", "NTHWC"),          # batches straddle events
        "sevir": (2, 10, 5, "sevir", "NTHWC"),    # the original SEVIR offsets / scales
        "nthw": (2, 13, 12, "01", "NTHW"),
    }.items():
        dl = object.__new__(SEVIRDataLoader)
        dl.data_types = ["vil"]
        dl.batch_size, dl.seq_len, dl.raw_seq_len, dl.stride = bs, seq_len, ev.shape[-1], stride
        dl.preprocess, dl.rescale_method, dl.layout, dl.downsample_dict = True, rescale, layout, None
        dl._load_event_batch = lambda event_idx, event_batch_size: [ev[event_idx:event_idx + event_batch_size]]
        n_batches = (ev.shape[0] * dl.num_seq_per_event) // bs
        out[f"{tag}_n"] = np.asarray(n_batches)
        for i in range(n_batches):
            out[f"{tag}_{i}"] = dl._idx_sample(i)["vil"]
    save("loader", **out)


def gen_catalog():
    """The catalog layer of the unmodified reference SEVIRDataLoader (sevir_dataloader.py:212-300, 360-390, 541-576): the
    constructor run on a synthetic catalog with `h5py.File` replaced by an in-memory mapping (h5py and the dataset are absent
    from the image) - the pandas filtering, sample order, shuffling, file bookkeeping and reads are the reference's own."""
    import catalog_cases as CC
    files = CC.catalog_files()
    h5 = types.ModuleType("h5py")
    h5.File = lambda path, mode="r": files[path.split("/data/", 1)[1]]
    sys.modules["h5py"] = h5
    from prediff.datasets.sevir.sevir_dataloader import SEVIRDataLoader
    out = {}
    for tag, (ckw, lkw) in CC.CASES.items():
        ckw = dict(ckw)
        dtypes = ckw.pop("data_types", ["vil"])
        dl = SEVIRDataLoader(data_types=dtypes, raw_seq_len=CC.T_RAW, sample_mode="sequent", sevir_catalog=CC.catalog_frame(),
                             sevir_data_dir="/data", preprocess=True, **ckw, **lkw)
        out[f"{tag}_files"] = np.asarray([str(v) for v in dl._samples["vil_filename"].values])
        out[f"{tag}_index"] = np.asarray(dl._samples["vil_index"].values, dtype=np.int64)
        out[f"{tag}_len"] = np.asarray(len(dl))
        for i in range(len(dl)):
            out[f"{tag}_b{i}"] = np.asarray(dl[i]["vil"])
        if ckw.get("shuffle"):   # every reset() reshuffles the current order with the same seed (:508-515, :273-274)
            dl.reset()
            out[f"{tag}_files_epoch2"] = np.asarray([str(v) for v in dl._samples["vil_filename"].values])
            out[f"{tag}_index_epoch2"] = np.asarray(dl._samples["vil_index"].values, dtype=np.int64)
        print(f"catalog {tag}: {len(dl._samples)} events, {len(dl)} batches")
    # the torch Dataset wrapper the Lightning datamodule hands to its DataLoaders (sevir_torch_wrap.py:73-163), aug_mode "0"
    sys.modules["lightning"].LightningDataModule = type("LightningDataModule", (), {})
    sys.modules["lightning"].seed_everything = lambda *a, **k: None
    from prediff.datasets.sevir.sevir_torch_wrap import SEVIRTorchDataset
    for tag, layout in (("thwc", "THWC"), ("cthw", "CTHW")):
        ds = SEVIRTorchDataset(seq_len=13, raw_seq_len=CC.T_RAW, stride=6, layout=layout, sevir_catalog=CC.catalog_frame(),
                               sevir_data_dir="/data", rescale_method="01", start_date=datetime.datetime(2019, 2, 1))
        out[f"ds_{tag}_len"] = np.asarray(len(ds))
        for i in range(len(ds)):
            out[f"ds_{tag}_{i}"] = ds[i].numpy()
        print(f"catalog dataset {layout}: {len(ds)} items of shape {tuple(ds[0].shape)}")
    save("catalog", **out)


def gen_skill():
    """SEVIRSkillScore of the unmodified reference (datasets/sevir/evaluation.py). `torchmetrics` and `h5py` are absent
    from the image; the metric only uses torchmetrics.Metric as a state container (add_state / reset), so an inert
    stand-in with exactly that is installed; all arithmetic (update / compute) is the reference's own."""
    import torch.nn as nn

    class Metric(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
            self._defaults = {}

        def add_state(self, name, default, dist_reduce_fx=None):
            self._defaults[name] = default.clone()
            setattr(self, name, default.clone())

        def reset(self):
            for k, v in self._defaults.items():
                setattr(self, k, v.clone())

    for name, attrs in (("torchmetrics", {"Metric": Metric}), ("h5py", {})):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    from prediff.datasets.sevir.evaluation import SEVIRSkillScore
    pred, target = skill_inputs()
    out = {}
    thr = (16, 74, 133, 160, 181, 219)
    for tag, pre in (("p1", "sevir"), ("p4", "sevir_pool4")):
        for mode in ("0", "1", "2"):
            sc = SEVIRSkillScore(layout="NTHWC", mode=mode, seq_len=6, preprocess_type=pre, threshold_list=thr,
                                 metrics_list=("csi", "bias", "sucr", "pod"), eps=1e-4)
            sc.update(pred.unsqueeze(-1), target.unsqueeze(-1))
            sc.update(pred.flip(0).unsqueeze(-1), target.unsqueeze(-1))   # a second batch accumulates
            out[f"{tag}_m{mode}_hits"] = sc.hits
            out[f"{tag}_m{mode}_misses"] = sc.misses
            out[f"{tag}_m{mode}_fas"] = sc.fas
            res = sc.compute()
            for m in ("csi", "bias", "sucr", "pod"):
                out[f"{tag}_m{mode}_{m}"] = np.stack([np.asarray(res[t][m], dtype=np.float64) for t in thr])
                out[f"{tag}_m{mode}_avg_{m}"] = np.asarray(res["avg"][m], dtype=np.float64)
    fin = torch.nan_to_num(pred) - torch.nan_to_num(target)
    out["mse_nonan"] = (fin.double() ** 2).mean()
    out["mae_nonan"] = fin.double().abs().mean()
    save("skill", **out)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)", flush=True)


@torch.no_grad()
def gen_schedule():
    import prediff.diffusion.utils as U
    ldm = ref_ldm(ref_unet(Wt.TINY_UNET), ref_vae(Wt.TINY_VAE), Wt.TINY_UNET, Wt.TINY_VAE)
    names = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
             "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
             "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]
    out = {n: getattr(ldm, n) for n in names}
    ts = U.make_ddim_timesteps("uniform", 50, 1000, verbose=False)
    for eta in (0.0, 1.0):
        sig, a, ap = U.make_ddim_sampling_parameters(ldm.alphas_cumprod.numpy(), ts, eta, verbose=False)
        out[f"ddim50_sigmas_eta{int(eta)}"] = sig
    out.update(ddim50_timesteps=ts, ddim50_alphas=a, ddim50_alphas_prev=ap)
    from prediff.models.utils import timestep_embedding
    out["timestep_embedding_256"] = timestep_embedding(torch.tensor([0, 1, 500, 981, 999]), 256)
    save("schedule", **out)


@torch.no_grad()
def gen_unet(tag, cfg, B, ts):
    m = ref_unet(cfg)
    x = inp(1234, B, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cond = inp(1235, B, cfg.t_in, cfg.h, cfg.w, cfg.c)
    t = torch.tensor(ts, dtype=torch.long)
    t0 = time.time()
    out = m(x, t, cond)
    print(f"unet_{tag}: forward {time.time() - t0:.1f}s, out std {out.std():.3f} absmax {out.abs().max():.3f}")
    save(f"unet_{tag}", t=t, out=out)


def ema_model(seed=5150):
    """A small module with a frozen parameter and dotted parameter names, seeded."""
    import torch.nn as nn
    g = torch.Generator().manual_seed(seed)
    m = nn.Sequential(nn.Linear(6, 5), nn.Sequential(nn.LayerNorm(5), nn.Linear(5, 3, bias=False)))
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape, generator=g))
    m[0].bias.requires_grad_(False)
    return m


def ema_drift(m, step):
    """Deterministic parameter change standing in for an optimizer step."""
    with torch.no_grad():
        for i, p in enumerate(m.parameters()):
            p.add_(0.01 * (step + 1) * torch.cos(torch.arange(p.numel(), dtype=torch.float32).reshape(p.shape) + i))


@torch.no_grad()
def gen_ema():
    """Shadow parameters of the unmodified reference LitEma (src/prediff/utils/ema.py) after 1, 2 and 12 updates
    (decay_t = min(decay, (1 + n) / (10 + n)) warm-up), and the store / copy_to / restore round trip."""
    from prediff.utils.ema import LitEma
    m = ema_model()
    ema = LitEma(m, decay=0.9)
    out = {"keys": np.array(sorted(k for k, _ in ema.named_buffers()))}
    for step in range(12):
        ema_drift(m, step)
        ema(m)
        if step + 1 in (1, 2, 12):
            for k, v in ema.named_buffers():
                out[f"s{step + 1}_{k}"] = v.clone()
    before = [p.clone() for p in m.parameters()]
    ema.store(m.parameters())
    ema.copy_to(m)
    out["swapped_0weight"] = m[0].weight.clone()
    out["swapped_0bias_frozen"] = m[0].bias.clone()
    ema.restore(m.parameters())
    assert all(torch.equal(a, b) for a, b in zip(before, m.parameters()))
    # every state_dict key (+ shape) of the reference LatentDiffusion(use_ema=True) on the tiny config: what a Lightning
    # checkpoint of the PreDiff module carries for this class
    ucfg, vcfg = Wt.TINY_UNET, Wt.TINY_VAE
    ldm = ref_ldm(ref_unet(ucfg), ref_vae(vcfg), ucfg, vcfg)
    ldm_ema = type(ldm)(torch_nn_module=ref_unet(ucfg), layout="NTHWC", data_shape=(ucfg.t_out, vcfg.h, vcfg.w, 1),
                        use_ema=True, latent_shape=(ucfg.t_out, ucfg.h, ucfg.w, ucfg.c), first_stage_model=ref_vae(vcfg),
                        cond_stage_model="__is_first_stage__")
    out["ldm_state_dict_keys"] = np.array([f"{k}|{','.join(str(d) for d in v.shape)}" for k, v in ldm_ema.state_dict().items()])
    save("ema", **out)


class DummyEps(torch.nn.Module):
    """A closed-form stand-in denoiser for the step-helper goldens (no weights)."""

    def forward(self, x, t, cond):
        return 0.5 * torch.tanh(x) + 0.1 * cond.mean(dim=1, keepdim=True)[:, :, :, :, :x.shape[-1]]


@torch.no_grad()
def gen_helpers():
    """predict_start_from_noise / q_posterior / p_mean_variance / aligned_mean of the unmodified reference
    LatentDiffusion (latent_diffusion.py:553-596) with a closed-form denoiser."""
    ldm = ref_ldm(DummyEps(), ref_vae(Wt.TINY_VAE), Wt.TINY_UNET, Wt.TINY_VAE)
    B = 3
    zt, zc, noise = inp(61, B, 6, 4, 4, 8), inp(62, B, 7, 4, 4, 8), inp(63, B, 6, 4, 4, 8)
    t = torch.tensor([0, 431, 999], dtype=torch.long)
    out = {"x0": ldm.predict_start_from_noise(zt, t, noise)}
    out["qp_mean"], out["qp_var"], out["qp_logvar"] = ldm.q_posterior(noise, zt, t)
    out["pm_mean"], out["pm_var"], out["pm_logvar"], out["pm_x0"] = ldm.p_mean_variance(zt, zc, t, clip_denoised=False, return_x0=True)
    out["pm_mean_clipped"] = ldm.p_mean_variance(zt, zc, t, clip_denoised=True)[0]
    ldm.set_alignment(lambda zt, t, zc=None, y=None, **kw: 0.2 * zt + kw["shift"])
    out["aligned"] = ldm.aligned_mean(zt, t, zc, None, out["pm_mean"], out["pm_logvar"], shift=0.05)
    save("helpers", **out)


LOSS_CASES = [("l2", dict(loss_type="l2")),
              ("l1w", dict(loss_type="l1", original_elbo_weight=0.3, l_simple_weight=0.7, logvar_init=0.5))]
# learn_logvar = True as in the shipped config (cfg.yaml:95), with per-timestep values standing in for a trained logvar
LEARNED_LOGVAR_CASE = ("learned", dict(loss_type="l2", learn_logvar=True, original_elbo_weight=0.1))


def learned_logvar_values():
    return torch.linspace(-0.5, 0.5, 1000)


@torch.no_grad()
def gen_losses():
    """Forward diffusion loss of the unmodified reference LatentDiffusion.p_losses (latent_diffusion.py:517-551) on the
    tiny UNet with injected t / noise (what validation_step evaluates through self(batch)), plus the lvlb_weights
    buffer (:270-277)."""
    ucfg, vcfg = Wt.TINY_UNET, Wt.TINY_VAE
    unet = ref_unet(ucfg)
    B = 3
    z = inp(881, B, ucfg.t_out, ucfg.h, ucfg.w, ucfg.c)
    zc = inp(882, B, ucfg.t_in, ucfg.h, ucfg.w, ucfg.c)
    noise = inp(883, B, ucfg.t_out, ucfg.h, ucfg.w, ucfg.c)
    t = torch.tensor([0, 431, 999], dtype=torch.long)
    out = {"t": t}
    for tag, kw in LOSS_CASES + [LEARNED_LOGVAR_CASE]:
        ldm = ref_ldm(unet, ref_vae(vcfg), ucfg, vcfg, **kw)   # .eval() -> 'val/' prefix; a fresh VAE per instance
        # (the reference replaces first_stage_model.train with an unbound function, so a VAE cannot be reused)
        if kw.get("learn_logvar"):
            ldm.logvar.data.copy_(learned_logvar_values())
        loss, d = ldm.p_losses(z, zc, t, noise=noise)
        out[f"{tag}_loss"] = loss
        for k, v in d.items():
            out[f"{tag}_{k.replace('/', '_')}"] = v
        out["lvlb_weights"] = ldm.lvlb_weights
        xn = ldm.q_sample(z, t, noise)
        out["x_noisy_probe"] = xn[:, 0, 0, 0, :8]
        print(f"losses {tag}: " + ", ".join(f"{k}={float(v):.6f}" for k, v in d.items()))
    save("losses", **out)


@torch.no_grad()
def gen_patterns():
    """Other cuboid patterns: single CuboidSelfAttentionLayer outputs of the unmodified reference
    class for shifted / padded / dilated / clipped cuboids under 'zeros' and 'ignore' padding, and forwards of the
    unmodified reference UNet built with non-axial block_attn_patterns."""
    import dataclasses
    from prediff.models.cuboid_transformer.cuboid_transformer import CuboidSelfAttentionLayer
    import pattern_cases as PC
    out = {}
    for tag, dims, C, heads, size, strat, shift, pad in PC.LAYER_CASES + PC.sweep_cases():
        m = CuboidSelfAttentionLayer(dim=C, num_heads=heads, cuboid_size=size, shift_size=shift, strategy=tuple(strat),
                                     padding_type=pad, qkv_bias=False, attn_drop=0.0, proj_drop=0.0,
                                     use_final_proj=True, norm_layer="layer_norm", use_global_vector=False,
                                     checkpoint_level=0, use_relative_pos=True).eval()
        sd = Wt.seeded_state_dict(PC.layer_spec(C, heads, size), PC.LAYER_SEED)
        res = m.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
        assert not res.unexpected_keys and res.missing_keys == ["relative_position_index"], res
        x = inp(PC.LAYER_SEED + 1, 2, *dims, C)
        out[f"layer_{tag}"] = m(x)
    for tag, pats, pad in PC.UNET_CASES + [("explicit", None, "ignore")]:
        if pats is None:
            cfg = dataclasses.replace(Wt.TINY_UNET, patterns=("explicit", "explicit"), padding_type=pad,
                                      explicit_layers=PC.EXPLICIT_LAYERS)
        else:
            cfg = dataclasses.replace(Wt.TINY_UNET, patterns=tuple(pats), padding_type=pad)
        m = ref_unet(cfg)
        x = inp(1234, 1, cfg.t_out, cfg.h, cfg.w, cfg.c)
        cond = inp(1235, 1, cfg.t_in, cfg.h, cfg.w, cfg.c)
        t0 = time.time()
        out[f"unet_{tag}"] = m(x, torch.tensor([500], dtype=torch.long), cond)
        print(f"patterns unet_{tag}: {time.time() - t0:.1f}s, out std {out[f'unet_{tag}'].std():.3f}")
    save("patterns", **out)


@torch.no_grad()
def gen_patterns_nearest():
    """padding_type='nearest' (models/utils.py:228-270): outputs of the unmodified CuboidSelfAttentionLayer and of the
    unmodified reference UNet (tiny config, a pattern that pads T = 13)."""
    import dataclasses
    from prediff.models.cuboid_transformer.cuboid_transformer import CuboidSelfAttentionLayer
    import pattern_cases as PC
    out = {}
    for tag, dims, C, heads, size, strat, shift, pad in PC.NEAREST_LAYER_CASES:
        m = CuboidSelfAttentionLayer(dim=C, num_heads=heads, cuboid_size=size, shift_size=shift, strategy=tuple(strat),
                                     padding_type=pad, qkv_bias=False, attn_drop=0.0, proj_drop=0.0,
                                     use_final_proj=True, norm_layer="layer_norm", use_global_vector=False,
                                     checkpoint_level=0, use_relative_pos=True).eval()
        sd = Wt.seeded_state_dict(PC.layer_spec(C, heads, size), PC.LAYER_SEED)
        res = m.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
        assert not res.unexpected_keys and res.missing_keys == ["relative_position_index"], res
        x = inp(PC.LAYER_SEED + 1, 2, *dims, C)
        out[f"layer_{tag}"] = m(x)
    for tag, pats, pad in PC.NEAREST_UNET_CASES:
        cfg = dataclasses.replace(Wt.TINY_UNET, patterns=tuple(pats), padding_type=pad)
        m = ref_unet(cfg)
        x = inp(1234, 1, cfg.t_out, cfg.h, cfg.w, cfg.c)
        cond = inp(1235, 1, cfg.t_in, cfg.h, cfg.w, cfg.c)
        out[f"unet_{tag}"] = m(x, torch.tensor([500], dtype=torch.long), cond)
        print(f"patterns_nearest unet_{tag}: out std {out[f'unet_{tag}'].std():.3f}")
    save("patterns_nearest", **out)


@torch.no_grad()
def gen_global_vectors():
    """Global vectors (cuboid_transformer.py:864-945, cuboid_transformer_unet.py:124-126, 432-434, 449-450, 489-490): outputs
    (x_out, new_global_vector) of the unmodified CuboidSelfAttentionLayer with use_global_vector=True, and forwards of the
    unmodified reference UNet built with num_global_vectors > 0."""
    import contextlib
    import dataclasses
    import io
    from prediff.models.cuboid_transformer.cuboid_transformer import CuboidSelfAttentionLayer
    import pattern_cases as PC
    out = {}
    for tag, dims, C, heads, size, strat, shift, pad, K, gsa in PC.GV_LAYER_CASES:
        m = CuboidSelfAttentionLayer(dim=C, num_heads=heads, cuboid_size=size, shift_size=shift, strategy=tuple(strat),
                                     padding_type=pad, qkv_bias=False, attn_drop=0.0, proj_drop=0.0,
                                     use_final_proj=True, norm_layer="layer_norm", use_global_vector=True,
                                     use_global_self_attn=gsa, separate_global_qkv=False, global_dim_ratio=1,
                                     checkpoint_level=0, use_relative_pos=True).eval()
        sd = Wt.seeded_state_dict(PC.gv_layer_spec(C, heads, size), PC.LAYER_SEED)
        res = m.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
        assert not res.unexpected_keys and res.missing_keys == ["relative_position_index"], res
        x = inp(PC.LAYER_SEED + 1, 2, *dims, C)
        g = inp(PC.LAYER_SEED + 2, 2, K, C)
        xo, go = m(x, g)
        out[f"layer_{tag}_x"], out[f"layer_{tag}_g"] = xo, go
        print(f"global_vectors layer_{tag}: x std {xo.std():.3f}, g std {go.std():.3f}")
    for tag, pats, pad, K, gffn, gsa in PC.GV_UNET_CASES:
        cfg = dataclasses.replace(Wt.TINY_UNET, patterns=tuple(pats), padding_type=pad, num_global_vectors=K,
                                  use_global_vector_ffn=gffn, use_global_self_attn=gsa)
        m = ref_unet(cfg)
        x = inp(1234, 2, cfg.t_out, cfg.h, cfg.w, cfg.c)
        cond = inp(1235, 2, cfg.t_in, cfg.h, cfg.w, cfg.c)
        with contextlib.redirect_stdout(io.StringIO()):   # the reference forward prints shapes on this path (:451-452)
            out[f"unet_{tag}"] = m(x, torch.tensor([500, 37], dtype=torch.long), cond)
        print(f"global_vectors unet_{tag}: out std {out[f'unet_{tag}'].std():.3f}")
    save("global_vectors", **out)


@torch.no_grad()
def gen_global_vectors_sep():
    """separate_global_qkv=True (cuboid_transformer.py:770-795, 866-891): the unmodified layer and UNet, as gen_global_vectors."""
    import contextlib
    import dataclasses
    import io
    from prediff.models.cuboid_transformer.cuboid_transformer import CuboidSelfAttentionLayer
    import pattern_cases as PC
    out = {}
    for tag, dims, C, heads, size, strat, shift, pad, K, gsa in PC.GV_SEP_LAYER_CASES:
        m = CuboidSelfAttentionLayer(dim=C, num_heads=heads, cuboid_size=size, shift_size=shift, strategy=tuple(strat),
                                     padding_type=pad, qkv_bias=False, attn_drop=0.0, proj_drop=0.0,
                                     use_final_proj=True, norm_layer="layer_norm", use_global_vector=True,
                                     use_global_self_attn=gsa, separate_global_qkv=True, global_dim_ratio=1,
                                     checkpoint_level=0, use_relative_pos=True).eval()
        spec = PC.gv_sep_layer_spec(C, heads, size, gsa)
        assert [k for k, _ in spec] == ["a." + k for k in m.state_dict() if k != "relative_position_index"], "registration order"
        sd = Wt.seeded_state_dict(spec, PC.LAYER_SEED)
        res = m.load_state_dict({k[2:]: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
        assert not res.unexpected_keys and res.missing_keys == ["relative_position_index"], res
        x = inp(PC.LAYER_SEED + 1, 2, *dims, C)
        g = inp(PC.LAYER_SEED + 2, 2, K, C)
        xo, go = m(x, g)
        out[f"layer_{tag}_x"], out[f"layer_{tag}_g"] = xo, go
        print(f"global_vectors_sep layer_{tag}: x std {xo.std():.3f}, g std {go.std():.3f}")
    for tag, pats, pad, K, gffn, gsa in PC.GV_SEP_UNET_CASES:
        cfg = dataclasses.replace(Wt.TINY_UNET, patterns=tuple(pats), padding_type=pad, num_global_vectors=K,
                                  use_global_vector_ffn=gffn, use_global_self_attn=gsa, separate_global_qkv=True)
        m = ref_unet(cfg)
        x = inp(1234, 2, cfg.t_out, cfg.h, cfg.w, cfg.c)
        cond = inp(1235, 2, cfg.t_in, cfg.h, cfg.w, cfg.c)
        with contextlib.redirect_stdout(io.StringIO()):
            out[f"unet_{tag}"] = m(x, torch.tensor([500, 37], dtype=torch.long), cond)
        print(f"global_vectors_sep unet_{tag}: out std {out[f'unet_{tag}'].std():.3f}")
    save("global_vectors_sep", **out)


@torch.no_grad()
def gen_vae(tag, cfg, N):
    m = ref_vae(cfg)
    x = inp(4321, N, 1, cfg.h, cfg.w, uniform=True)
    post = m.encode(x)
    moments = post.parameters
    z = post.mode()
    dec = m.decode(z)
    print(f"vae_{tag}: moments std {moments.std():.3f}, dec std {dec.std():.3f}")
    save(f"vae_{tag}", moments=moments, dec=dec)


@torch.no_grad()
def gen_loop_tiny():
    """Reference p_sample_loop (DDPM ancestral) and sample() on the tiny config, RNG injected."""
    import prediff.diffusion.latent_diffusion as LD
    ucfg, vcfg = Wt.TINY_UNET, Wt.TINY_VAE
    unet, vae = ref_unet(ucfg), ref_vae(vcfg)
    ldm = ref_ldm(unet, vae, ucfg, vcfg)
    B, n_steps = 2, 4
    zT = inp(777, B, ucfg.t_out, ucfg.h, ucfg.w, ucfg.c)
    cond = inp(778, B, ucfg.t_in, ucfg.h, ucfg.w, ucfg.c)
    noise = inp(779, n_steps, B, ucfg.t_out, ucfg.h, ucfg.w, ucfg.c)
    calls = {"k": 0}

    def fake_noise_like(shape, device):
        k = calls["k"]
        calls["k"] += 1
        assert tuple(shape) == tuple(noise[k].shape)
        return noise[k].clone()

    orig = LD.noise_like
    LD.noise_like = fake_noise_like
    try:
        z0, inter = ldm.p_sample_loop(cond=cond, shape=tuple(zT.shape), x_T=zT.clone(), timesteps=n_steps,
                                      return_intermediates=True, log_every_t=1)
        assert calls["k"] == n_steps
        # one step at a large timestep (p_sample directly)
        calls["k"] = 0
        t = torch.full((B,), 900, dtype=torch.long)
        z_step = ldm.p_sample(zt=zT.clone(), zc=cond, t=t)
        # full sample(): encode context -> loop -> decode
        calls["k"] = 0
        y = inp(780, B, ucfg.t_in, vcfg.h, vcfg.w, 1, uniform=True)
        dec = ldm.sample(cond={"y": y}, batch_size=B, x_T=zT.clone(), timesteps=n_steps)
        zc = ldm.cond_stage_forward({"y": y})
    finally:
        LD.noise_like = orig
    assert len(inter) == n_steps + 1 and torch.equal(inter[-1], z0) and torch.equal(inter[0], zT)
    save("loop_tiny", z0=z0, inter1=inter[1], z_step900=z_step, sample_dec=dec, sample_zc=zc)


@torch.no_grad()
def gen_loop_shorten():
    """Reference p_sample_loop with num_timesteps_cond = 4 (shorten_cond_schedule, latent_diffusion.py:153-157, 295-299,
    665-667): the context latents are re-noised before every step. Tiny config; both RNG streams injected (noise_like for
    the ancestral noise, torch.randn_like for the context noise)."""
    import prediff.diffusion.latent_diffusion as LD
    ucfg, vcfg = Wt.TINY_UNET, Wt.TINY_VAE
    ldm = ref_ldm(ref_unet(ucfg), ref_vae(vcfg), ucfg, vcfg, num_timesteps_cond=4)
    assert ldm.shorten_cond_schedule
    B, n_steps = 2, 4
    zT = inp(777, B, ucfg.t_out, ucfg.h, ucfg.w, ucfg.c)
    cond = inp(778, B, ucfg.t_in, ucfg.h, ucfg.w, ucfg.c)
    noise = inp(779, n_steps, B, ucfg.t_out, ucfg.h, ucfg.w, ucfg.c)
    cnoise = inp(781, n_steps, B, ucfg.t_in, ucfg.h, ucfg.w, ucfg.c)
    calls = {"k": 0, "c": 0}

    def fake_noise_like(shape, device):
        k = calls["k"]
        calls["k"] += 1
        return noise[k].clone()

    def fake_randn_like(x, **kw):
        c = calls["c"]
        calls["c"] += 1
        assert tuple(x.shape) == tuple(cnoise[c].shape)
        return cnoise[c].clone()

    orig, orig_rl = LD.noise_like, torch.randn_like
    LD.noise_like = fake_noise_like
    torch.randn_like = fake_randn_like
    try:
        z0 = ldm.p_sample_loop(cond=cond, shape=tuple(zT.shape), x_T=zT.clone(), timesteps=n_steps)
    finally:
        LD.noise_like = orig
        torch.randn_like = orig_rl
    assert calls["k"] == n_steps and calls["c"] == n_steps
    save("loop_shorten", z0=z0, cond_ids=ldm.cond_ids)


@torch.no_grad()
def gen_loop_extra():
    """Round-2 pins: (1) clip_denoised=True through the reference's p_sample / p_sample_loop (latent_diffusion.py:580-581),
    tiny config, RNG injected; (2) the reference's own sample() at the SHIPPED sizes: 7 context frames 128x128 -> encode
    -> 4 ancestral steps -> decode; (3) the same full-size encode / decode around the S6 50-step DDIM (reference UNet +
    the reference's DDIM helpers), i.e. what `sample(sampler="ddim")` must reproduce end to end."""
    import prediff.diffusion.latent_diffusion as LD
    import prediff.diffusion.utils as U
    out = {}
    ucfg, vcfg = Wt.TINY_UNET, Wt.TINY_VAE
    ldm = ref_ldm(ref_unet(ucfg), ref_vae(vcfg), ucfg, vcfg)
    B, n_steps = 2, 4
    zT = inp(777, B, ucfg.t_out, ucfg.h, ucfg.w, ucfg.c)
    cond = inp(778, B, ucfg.t_in, ucfg.h, ucfg.w, ucfg.c)
    noise = inp(779, n_steps, B, ucfg.t_out, ucfg.h, ucfg.w, ucfg.c)
    calls = {"k": 0, "noise": noise}

    def fake_noise_like(shape, device):
        k = calls["k"]
        calls["k"] += 1
        assert tuple(shape) == tuple(calls["noise"][k].shape)
        return calls["noise"][k].clone()

    orig = LD.noise_like
    LD.noise_like = fake_noise_like
    try:
        t = torch.full((B,), 100, dtype=torch.long)
        # t = 100 and a half-scale z_t: about a quarter of the z_0 estimate is clamped, the rest passes through
        out["z_step100_clip"] = ldm.p_sample(zt=zT.clone() * 0.5, zc=cond, t=t, clip_denoised=True)
        calls["k"] = 0
        out["z_step100_clip_x0"] = ldm.p_sample(zt=zT.clone() * 0.5, zc=cond, t=t, clip_denoised=True, return_x0=True)[1]
        calls["k"] = 0
        ldm.clip_denoised = True
        out["z0_clip"] = ldm.p_sample_loop(cond=cond, shape=tuple(zT.shape), x_T=zT.clone(), timesteps=n_steps)
        ldm.clip_denoised = False
        frac = ((out["z_step100_clip_x0"].abs() == 1.0).float().mean()).item()
        print(f"loop_extra: clip_denoised clamps {100 * frac:.1f}% of the z_0 estimate at t=100")
        # ---- shipped sizes, batch 1 ----
        fu, fv = Wt.UNetConfig(), Wt.VAEConfig()
        unet_f = ref_unet(fu)
        ldm_f = ref_ldm(unet_f, ref_vae(fv), fu, fv)
        y = inp(880, 1, fu.t_in, fv.h, fv.w, 1, uniform=True)
        zT_f = inp(881, 1, fu.t_out, fu.h, fu.w, fu.c)
        calls["noise"] = inp(882, n_steps, 1, fu.t_out, fu.h, fu.w, fu.c)
        calls["k"] = 0
        t0 = time.time()
        out["full_sample_ddpm4"] = ldm_f.sample(cond={"y": y}, batch_size=1, x_T=zT_f.clone(), timesteps=n_steps)
        zc = ldm_f.cond_stage_forward({"y": y})
        out["full_zc"] = zc
        print(f"loop_extra: full-size sample() with 4 ancestral steps: {time.time() - t0:.0f}s", flush=True)
    finally:
        LD.noise_like = orig
    betas = U.make_beta_schedule("linear", 1000, linear_start=1e-4, linear_end=2e-2)
    ac = np.cumprod(1.0 - betas, axis=0).astype(np.float32)
    ts = U.make_ddim_timesteps("uniform", 50, 1000, verbose=False)
    sig, a, ap = U.make_ddim_sampling_parameters(ac, ts, 0.0, verbose=False)
    z = zT_f.clone()
    t0 = time.time()
    for k, i in enumerate(reversed(range(len(ts)))):
        t = torch.full((1,), int(ts[i]), dtype=torch.long)
        eps = unet_f(z, t, zc)
        z0 = (z - float(np.sqrt(1.0 - a[i])) * eps) / float(np.sqrt(a[i]))
        z = float(np.sqrt(ap[i])) * z0 + float(np.sqrt(1.0 - ap[i] - sig[i] ** 2)) * eps
        if k % 10 == 0:
            print(f"loop_extra: full ddim step {k + 1}/50 ({time.time() - t0:.0f}s)", flush=True)
    out["full_ddim50_z0"] = z
    out["full_sample_ddim50"] = ldm_f.decode_first_stage(z)
    save("loop_extra", **out)


@torch.no_grad()
def gen_ddim(tag, cfg, B, n_steps):
    """S6 DDIM (eta = 0) with the reference UNet module and the reference's DDIM helper functions."""
    import prediff.diffusion.utils as U
    m = ref_unet(cfg)
    betas = U.make_beta_schedule("linear", 1000, linear_start=1e-4, linear_end=2e-2)
    ac = np.cumprod(1.0 - betas, axis=0).astype(np.float32)  # the registered fp32 buffer
    ts = U.make_ddim_timesteps("uniform", n_steps, 1000, verbose=False)
    sig, a, ap = U.make_ddim_sampling_parameters(ac, ts, 0.0, verbose=False)
    z = inp(4242, B, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cond = inp(4243, B, cfg.t_in, cfg.h, cfg.w, cfg.c)
    t0 = time.time()
    snaps = {}
    for k, i in enumerate(reversed(range(len(ts)))):
        t = torch.full((B,), int(ts[i]), dtype=torch.long)
        eps = m(z, t, cond)
        z0 = (z - float(np.sqrt(1.0 - a[i])) * eps) / float(np.sqrt(a[i]))
        z = float(np.sqrt(ap[i])) * z0 + float(np.sqrt(1.0 - ap[i] - sig[i] ** 2)) * eps
        if k + 1 in (1, 25) and tag == "full":
            snaps[f"z_after_{k + 1}"] = z.clone()
        if k % 10 == 0:
            print(f"ddim_{tag}: step {k + 1}/{len(ts)} ({time.time() - t0:.0f}s)", flush=True)
    save(f"ddim_{tag}", z0=z, **snaps)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="schedule,unet_tiny,unet_full,vae_tiny,vae_full,loop_tiny,ddim_tiny")
    args = ap.parse_args()
    install_stubs()
    torch.manual_seed(0)
    todo = set(args.only.split(","))
    FULL_U, FULL_V = Wt.UNetConfig(), Wt.VAEConfig()
    if "schedule" in todo:
        gen_schedule()
    if "unet_tiny" in todo:
        gen_unet("tiny", Wt.TINY_UNET, 2, [500, 37])
    if "vae_tiny" in todo:
        gen_vae("tiny", Wt.TINY_VAE, 2)
    if "loop_tiny" in todo:
        gen_loop_tiny()
    if "ddim_tiny" in todo:
        gen_ddim("tiny", Wt.TINY_UNET, 2, 50)
    if "unet_full" in todo:
        gen_unet("full", FULL_U, 1, [500])
    if "vae_full" in todo:
        gen_vae("full", FULL_V, 1)
    if "ka" in todo:
        gen_ka()
    if "skill" in todo:
        gen_skill()
    if "loader" in todo:
        gen_loader()
    if "catalog" in todo:
        gen_catalog()
    if "patterns" in todo:
        gen_patterns()
    if "patterns_nearest" in todo:
        gen_patterns_nearest()
    if "global_vectors" in todo:
        gen_global_vectors()
    if "global_vectors_sep" in todo:
        gen_global_vectors_sep()
    if "losses" in todo:
        gen_losses()
    if "ema" in todo:
        gen_ema()
    if "helpers" in todo:
        gen_helpers()
    if "loop_extra" in todo:
        gen_loop_extra()
    if "loop_shorten" in todo:
        gen_loop_shorten()
    if "ddim_full" in todo:
        gen_ddim("full", FULL_U, 4, 50)
