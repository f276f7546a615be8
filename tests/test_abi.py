"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/prediff_b200.h declares, declares the reference's state_dict keys, reproduces the reference schedule
buffers bit-exactly, and refuses to compute without an sm_100 device (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from prediff_b200 import _lib as L
from prediff_b200 import weights as Wt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "prediff_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert b"sm_100a" in lib.pd_version()


def test_unet_and_vae_weight_specs_match_reference_keys():
    from prediff_b200.unet import CuboidTransformerUNet
    from prediff_b200.vae import AutoencoderKL
    for cfg in (Wt.TINY_UNET, Wt.UNetConfig()):
        m = CuboidTransformerUNet([cfg.t_in, cfg.h, cfg.w, cfg.c], [cfg.t_out, cfg.h, cfg.w, cfg.c],
                                  base_units=cfg.base_units, depth=list(cfg.depth), num_heads=cfg.num_heads)
        spec = [(n, tuple(s)) for n, s in Wt.unet_param_spec(cfg)]
        assert m.weight_spec_from_library() == spec
        assert [(n, tuple(p.shape)) for n, p in m.named_parameters()] == spec
    # global vectors (cuboid_transformer_unet.py:124-126, 167-191; cuboid_transformer.py:777-810, 1054-1068): the extra keys, in
    # the reference's registration order (pinned against the reference's state_dict by tests/golden/gen_golden.py::ref_unet)
    import dataclasses
    for gffn, pats, sep, gsa in ((True, ("axial", "axial"), False, True), (False, ("video_swin_2x8", "spatial_lg_4"), False, True),
                                 (False, ("axial", "axial"), True, True), (True, ("axial", "axial"), True, False)):
        cfg = dataclasses.replace(Wt.TINY_UNET, num_global_vectors=4, use_global_vector_ffn=gffn, patterns=pats,
                                  separate_global_qkv=sep, use_global_self_attn=gsa)
        m = CuboidTransformerUNet([cfg.t_in, cfg.h, cfg.w, cfg.c], [cfg.t_out, cfg.h, cfg.w, cfg.c], base_units=cfg.base_units,
                                  depth=list(cfg.depth), num_heads=cfg.num_heads, block_attn_patterns=list(pats),
                                  num_global_vectors=4, use_global_vector_ffn=gffn, use_global_self_attn=gsa,
                                  separate_global_qkv=sep)
        assert any(".g2g_global_qkv_net." in n for n, _ in Wt.unet_param_spec(cfg)) == (sep and gsa)
        assert any(".global_qkv." in n for n, _ in Wt.unet_param_spec(cfg)) == (not sep)
        spec = [(n, tuple(s)) for n, s in Wt.unet_param_spec(cfg)]
        assert spec[0] == ("init_global_vectors", (4, 64))
        assert m.weight_spec_from_library() == spec
        assert [(n, tuple(p.shape)) for n, p in m.named_parameters()] == spec
        assert any(".global_ffn_l." in n for n, _ in spec) == gffn
    full = Wt.UNetConfig()
    assert sum(int(np.prod(s)) for _, s in Wt.unet_param_spec(full)) == 136817538
    for cfg in (Wt.TINY_VAE, Wt.VAEConfig()):
        m = AutoencoderKL(block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block,
                          latent_channels=cfg.latent_channels, sample_size=(cfg.h, cfg.w))
        spec = [(n, tuple(s)) for n, s in Wt.vae_param_spec(cfg)]
        assert m.weight_spec_from_library() == spec
    assert sum(int(np.prod(s)) for _, s in Wt.vae_param_spec(Wt.VAEConfig())) == 84499393


def test_ka_weight_spec_matches_reference_keys():
    from prediff_b200.alignment import NoisyCuboidTransformerEncoder, SEVIRAvgIntensityAlignment
    cfg = Wt.KAConfig()
    al = SEVIRAvgIntensityAlignment(guide_scale=50.0, model_args=dict(
        input_shape=[cfg.t, cfg.h, cfg.w, cfg.c], base_units=cfg.base_units, depth=list(cfg.depth), block_attn_patterns="axial",
        num_heads=cfg.num_heads, pool="attention", readout_seq=True, out_len=cfg.t))
    spec = [(n, tuple(s)) for n, s in Wt.ka_param_spec(cfg)]
    assert al.model.weight_spec_from_library() == spec
    assert [(n, tuple(p.shape)) for n, p in al.model.named_parameters()] == spec
    assert sum(int(np.prod(s)) for _, s in spec) == 8937033
    with pytest.raises(NotImplementedError):
        NoisyCuboidTransformerEncoder([6, 16, 16, 64], pool="adaptive")
    with pytest.raises(NotImplementedError):
        NoisyCuboidTransformerEncoder([6, 16, 16, 64], block_attn_patterns="video_swin_2x4")


def test_state_dict_has_reference_keys_including_derived_buffers():
    from prediff_b200.unet import CuboidTransformerUNet
    cfg = Wt.TINY_UNET
    m = CuboidTransformerUNet([cfg.t_in, cfg.h, cfg.w, cfg.c], [cfg.t_out, cfg.h, cfg.w, cfg.c], base_units=64, depth=[1, 1])
    sd = m.state_dict()
    assert "down_self_blocks.0.0.attn_l.2.qkv.weight" in sd and tuple(sd["down_self_blocks.0.0.attn_l.2.qkv.weight"].shape) == (192, 64)
    idx = sd["down_self_blocks.1.0.attn_l.1.relative_position_index"]
    assert idx.dtype == torch.int64 and tuple(idx.shape) == (8, 8) and idx[0, 0] == 7 and idx[7, 0] == 14
    # loading a dict that carries the buffers (as a real pretrained .pt does) works with strict=True
    m.load_state_dict(sd, strict=True)


def test_schedule_through_c_abi_is_bit_exact_vs_reference():
    g = np.load(os.path.join(ROOT, "tests", "golden", "schedule.npz"))
    h = ctypes.c_void_p()
    L.check(L.lib().pd_sampler_create(1000, ctypes.c_double(1e-4), ctypes.c_double(2e-2), ctypes.byref(h)))
    for name in ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                 "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                 "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]:
        buf = np.empty(1000, dtype=np.float32)
        L.check(L.lib().pd_sampler_get_buffer(h, name.encode(), buf.ctypes.data_as(ctypes.c_void_p)))
        assert np.array_equal(buf, g[name]), name
    with pytest.raises(L.PDError):
        L.check(L.lib().pd_sampler_get_buffer(h, b"no_such_buffer", buf.ctypes.data_as(ctypes.c_void_p)))
    L.lib().pd_sampler_destroy(h)


def test_unsupported_configs_fail_loudly():
    from prediff_b200.unet import CuboidTransformerUNet
    with pytest.raises(NotImplementedError):
        CuboidTransformerUNet([7, 16, 16, 64], [6, 16, 16, 64], block_attn_patterns="video_swin_3x5")  # not registered
    with pytest.raises(NotImplementedError):
        CuboidTransformerUNet([7, 16, 16, 64], [6, 16, 16, 64], padding_type="reflect")   # not one of the reference's three
    for kw in (dict(num_global_vectors=8, separate_global_qkv=True, global_dim_ratio=2), dict(num_global_vectors=8, global_dim_ratio=2),
               dict(num_global_vectors=33), dict(num_global_vectors=8, precision="tf32")):
        with pytest.raises(NotImplementedError):
            CuboidTransformerUNet([7, 16, 16, 64], [6, 16, 16, 64], **kw)
    with pytest.raises(NotImplementedError):
        CuboidTransformerUNet([7, 16, 16, 64], [6, 16, 16, 64], depth=[2, 2, 2])
    with pytest.raises(L.PDError):  # rejected by the C++ validate(): 24x24 latents do not tile
        CuboidTransformerUNet([7, 24, 24, 64], [6, 24, 24, 64], base_units=64, depth=[1, 1]).weight_spec_from_library()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_error_not_fallback():
    from prediff_b200.unet import CuboidTransformerUNet
    cfg = Wt.TINY_UNET
    m = CuboidTransformerUNet([cfg.t_in, cfg.h, cfg.w, cfg.c], [cfg.t_out, cfg.h, cfg.w, cfg.c], base_units=64, depth=[1, 1])
    with pytest.raises(L.PDError):
        m(torch.zeros(1, 6, 16, 16, 64), torch.zeros(1, dtype=torch.long), torch.zeros(1, 7, 16, 16, 64))
    with pytest.raises(L.PDError):
        L.init()


def test_lvlb_weights_through_c_abi_bit_exact_vs_reference():
    """The non-persistent lvlb_weights buffer (latent_diffusion.py:270-277) computed by the C++ sampler == the
    reference's fp32 tensor arithmetic (tests/golden/losses.npz)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "losses.npz"))
    h = ctypes.c_void_p()
    L.check(L.lib().pd_sampler_create(1000, ctypes.c_double(1e-4), ctypes.c_double(2e-2), ctypes.byref(h)))
    buf = np.empty(1000, dtype=np.float32)
    L.check(L.lib().pd_sampler_get_buffer(h, b"lvlb_weights", buf.ctypes.data_as(ctypes.c_void_p)))
    L.lib().pd_sampler_destroy(h)
    assert np.array_equal(buf, g["lvlb_weights"])


def test_pattern_and_loss_arguments_fail_loudly():
    """Bad cuboid-layer specs / loss options are rejected with an error (no silent default, no fallback)."""
    import torch.nn as nn
    from prediff_b200.diffusion import LatentDiffusion
    i3 = ctypes.c_int32 * 3
    meta = (ctypes.c_int32 * 12)()
    lib = L.lib()
    # padding_type is 0 / 1 / 2; cuboid sizes must be >= 1; strategies are 0 / 1
    for size, strat, shift, pad in (((4, 4, 4), (0, 0, 0), (0, 0, 0), 3), ((0, 4, 4), (0, 0, 0), (0, 0, 0), 0),
                                    ((4, 4, 4), (0, 2, 0), (0, 0, 0), 0), ((4, 4, 4), (0, 0, 0), (-1, 0, 0), 0)):
        rc = lib.pd_cuboid_tables(13, 16, 16, i3(*size), i3(*strat), i3(*shift), pad, meta, None, None, None, ctypes.c_int64(0))
        assert rc < 0 and lib.pd_last_error()
    # output arrays too small for the tables
    buf = np.zeros(8, np.int32)
    rc = lib.pd_cuboid_tables(13, 16, 16, i3(4, 4, 4), i3(0, 0, 0), i3(0, 0, 0), 0, meta, buf.ctypes.data_as(ctypes.c_void_p),
                              None, None, ctypes.c_int64(8))
    assert rc < 0
    # pd_unet_create_ex: number of attention layers per block outside 1..PD_MAX_ATTN_LAYERS
    from prediff_b200.unet import _CUnetConfig, _CUnetPattern
    cc = _CUnetConfig(7, 6, 16, 16, 64, 64, (ctypes.c_int32 * 2)(1, 1), 4, 2)
    pt = _CUnetPattern()
    pt.n_layers[0], pt.n_layers[1] = 0, 9
    h = ctypes.c_void_p()
    assert lib.pd_unet_create_ex(ctypes.byref(cc), ctypes.byref(pt), ctypes.byref(h)) < 0 and not h.value

    class Eps(nn.Module):
        def forward(self, x, t, c):
            return x

    for kw in (dict(loss_type="huber"), dict(parameterization="x0"), dict(scale_by_std=True),
               dict(cond_stage_forward="encode_first")):
        with pytest.raises(NotImplementedError):
            LatentDiffusion(torch_nn_module=Eps(), **kw)
    assert LatentDiffusion(torch_nn_module=Eps(), clip_denoised=True).clip_denoised   # built: clamp inside the update kernel
    assert LatentDiffusion(torch_nn_module=Eps(), num_timesteps_cond=4).shorten_cond_schedule   # built: sampling loop only
    ld = LatentDiffusion(torch_nn_module=Eps(), loss_type="l1", original_elbo_weight=0.5)
    assert ld.lvlb_weights.shape == (1000,) and "lvlb_weights" not in ld.state_dict()   # non-persistent, as in the reference
    assert tuple(ld.logvar.shape) == (1000,) and ld.loss_mean_dim == (1, 2, 3, 4)


def test_shape_properties_of_the_reference_classes():
    from prediff_b200.unet import CuboidTransformerUNet
    from prediff_b200.vae import AutoencoderKL
    m = CuboidTransformerUNet([7, 16, 16, 64], [6, 16, 16, 64], base_units=256, depth=[4, 4])
    assert m.data_shape == (13, 16, 16, 65)                                   # cuboid_transformer_unet.py:377-384
    assert m.mem_shapes == [(13, 16, 16, 256), (13, 8, 8, 512)]               # :386-404
    v = AutoencoderKL()
    v.enable_slicing()
    assert v.use_slicing
    v.disable_slicing()
    assert not v.use_slicing
