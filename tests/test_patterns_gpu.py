"""GPU parity for the non-axial cuboid patterns, through the C ABI:
  - pd_op_cuboid_attention (general cuboid attention kernel) vs the reference-pinned oracle core on the same bf16 q|k|v,
    for shifted / padded / dilated / clipped / multi-chunk cuboids, head dims 16..128, 'zeros', 'ignore' and 'nearest' padding;
  - the general kernel vs the axial fast-path kernel on axial layers;
  - CuboidTransformerUNet built with non-axial block_attn_patterns vs goldens of the unmodified reference UNet and vs
    the oracle at the shipped widths (256 / 512: fused FFN and cluster-LayerNorm paths with 1, 2 and 5 layers per block).
Tolerances: bf16 operands / probabilities, fp32 accumulate - max-abs error <= 1e-2 of the output's abs-max for the op,
rel-RMS <= 1.5e-2 and max-abs <= 4e-2 for a UNet forward (same bars as tests/test_unet_gpu.py)."""
import ctypes
import dataclasses
import os

import numpy as np
import pytest
import torch

from oracle import prediff_oracle as O
from prediff_b200 import _lib as L
from prediff_b200 import weights as Wt
from prediff_b200.unet import CuboidTransformerUNet
from tests.golden import pattern_cases as PC
from tests.golden.gen_golden import UNET_SEED, inp

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "patterns.npz"))
GN = np.load(os.path.join(os.path.dirname(__file__), "golden", "patterns_nearest.npz"))   # padding_type='nearest'
REL_RMS_TOL, MAX_TOL = 1.5e-2, 4e-2
I3 = ctypes.c_int32 * 3


def errs(out, ref):
    out = out.detach().double().cpu()
    ref = torch.as_tensor(np.asarray(ref)).double()
    rel_rms = ((out - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    mx = ((out - ref).abs().max() / ref.abs().max()).item()
    return rel_rms, mx


def run_op(qkv, table, dims, C, heads, size, strat, shift, pad, impl=0):
    """impl: 0 = the kernel the models would pick, 1 = warp-level mma.sync kernel, 2 = tcgen05 tile kernel."""
    B = qkv.shape[0]
    out = torch.full((B, *dims, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(L.lib().pd_op_cuboid_attention_impl(L.ptr(qkv), L.ptr(table), L.ptr(out), B, *dims, C, heads, I3(*size),
                                                I3(*[0 if s == "l" else 1 for s in strat]), I3(*shift),
                                                {"zeros": 0, "ignore": 1, "nearest": 2}[pad], impl, L.stream_ptr()))
    torch.cuda.synchronize()
    return out


# tcgen05 tile kernel (csrc/attention_tc.cu; VERDICT r01 +J1): head dims 64 / 128, every kind of cuboid geometry
TC_CASES = [
    ((13, 16, 16), 4, 64, (2, 8, 8), "lll", (0, 0, 0), "zeros"),      # video_swin_2x8 level 0: volume 128, one chunk, T 13 -> 14
    ((13, 16, 16), 4, 64, (2, 8, 8), "lll", (1, 4, 4), "ignore"),     # ... shifted windows + 'ignore' padding mask
    ((13, 8, 8), 4, 128, (2, 8, 8), "lll", (1, 4, 4), "ignore"),      # level 1 (head dim 128: two 64-channel slabs)
    ((13, 8, 8), 4, 128, (2, 8, 8), "lll", (0, 0, 0), "zeros"),
    ((13, 16, 16), 4, 64, (1, 16, 16), "lll", (0, 0, 0), "zeros"),    # divided_st plane: volume 256 = 2 tiles x 2 chunks
    ((13, 16, 16), 4, 64, (13, 16, 16), "lll", (0, 0, 0), "zeros"),   # full, level 0: 26 tiles x 26 chunks, 24 025-row table
    ((13, 8, 8), 4, 128, (13, 8, 8), "lll", (0, 0, 0), "zeros"),      # full, level 1: 832 = 6.5 tiles (ragged tile + chunk)
    ((13, 16, 16), 4, 64, (4, 4, 4), "lll", (2, 2, 2), "ignore"),     # volume 64 < one tile: half-empty tile, masked tail
    ((13, 16, 16), 2, 64, (4, 8, 8), "ldd", (0, 0, 0), "zeros"),      # dilated gather, volume 256, T 13 -> 16 zero slots
    ((6, 7, 9), 2, 64, (4, 7, 9), "lll", (2, 3, 4), "ignore"),        # ragged grid, volume 252, shifted
    ((13, 8, 8), 4, 128, (2, 8, 8), "lll", (1, 4, 4), "nearest"),     # 'nearest': T 13 -> 14 by resampling, shifted
    ((13, 16, 16), 4, 64, (4, 8, 8), "ldd", (0, 0, 0), "nearest"),    # 'nearest' + dilated gather, T 13 -> 16
]


@pytest.mark.parametrize("case", TC_CASES, ids=[f"{c[0]}-hd{c[2]}-{c[3]}-{c[4]}-{c[5]}-{c[6]}" for c in TC_CASES])
def test_cuboid_attention_tcgen05_vs_oracle_and_mma_sync(case):
    dims, heads, hd, size, strat, shift, pad = case
    C, B = heads * hd, 2
    g = torch.Generator().manual_seed(11)
    qkv = torch.randn(B, *dims, 3 * C, generator=g).bfloat16()
    n_rel = (2 * size[0] - 1) * (2 * size[1] - 1) * (2 * size[2] - 1)
    table = 0.5 * torch.randn(n_rel, heads, generator=g)
    out = run_op(qkv.cuda(), table.cuda(), dims, C, heads, size, strat, shift, pad, impl=2)
    ref = O.cuboid_attention_core(qkv.float(), table, heads, size, tuple(strat), shift, pad)
    assert torch.isfinite(out.float()).all()   # every real token written exactly once (output pre-filled with NaN)
    rel_rms, mx = errs(out.float(), ref)
    old = run_op(qkv.cuda(), table.cuda(), dims, C, heads, size, strat, shift, pad, impl=1)
    r2, m2 = errs(out.float(), old.float().cpu())
    print(f"tcgen05 cuboid attention {case}: vs oracle rel_rms={rel_rms:.2e} max={mx:.2e}; vs mma.sync {r2:.2e} / {m2:.2e}")
    assert rel_rms < 6e-3 and mx < 1e-2, (rel_rms, mx)
    assert r2 < 6e-3 and m2 < 1.5e-2, (r2, m2)
    # deterministic
    assert torch.equal(out, run_op(qkv.cuda(), table.cuda(), dims, C, heads, size, strat, shift, pad, impl=2))


# (dims, heads, head_dim, cuboid, strategy, shift, padding)
OP_CASES = [
    ((13, 16, 16), 4, 64, (4, 4, 4), "lll", (2, 2, 2), "zeros"),      # video_swin_4x4, T padded 13 -> 16, shifted
    ((13, 16, 16), 4, 64, (4, 4, 4), "lll", (2, 2, 2), "ignore"),
    ((13, 8, 8), 4, 128, (2, 8, 8), "lll", (1, 4, 4), "ignore"),      # video_swin_2x8 at level 1: volume 128, 2 chunks
    ((13, 8, 8), 4, 128, (2, 8, 8), "lll", (0, 0, 0), "zeros"),
    ((13, 16, 16), 4, 64, (1, 16, 16), "lll", (0, 0, 0), "zeros"),    # divided_st: volume 256, 4 q-tiles x 4 chunks
    ((13, 8, 8), 4, 128, (1, 4, 4), "ddd", (0, 0, 0), "zeros"),       # spatial_lg_4 dilated layer
    ((13, 8, 8), 4, 32, (1, 4, 1), "ddd", (0, 0, 0), "ignore"),       # axial_space_dilate_2
    ((6, 7, 9), 2, 16, (4, 3, 4), "ldl", (2, 1, 2), "zeros"),         # ragged everything
    ((6, 7, 9), 2, 16, (4, 3, 4), "ldl", (2, 1, 2), "ignore"),
    ((5, 6, 6), 2, 32, (8, 4, 8), "lll", (4, 2, 4), "ignore"),        # cuboid clipped to the data (index-buffer quirk)
    ((13, 16, 16), 4, 64, (13, 16, 16), "lll", (0, 0, 0), "zeros"),   # full attention: one 3328-token cuboid, 52 chunks
    ((13, 16, 16), 4, 16, (13, 1, 1), "lll", (0, 0, 0), "zeros"),     # axial layer through the general kernel
    ((13, 16, 16), 4, 64, (4, 4, 4), "lll", (2, 2, 2), "nearest"),    # 'nearest' padding (models/utils.py:228-270)
    ((6, 7, 9), 2, 16, (4, 3, 4), "ldl", (2, 1, 2), "nearest"),
    ((5, 7, 9), 2, 32, (3, 4, 5), "lld", (1, 2, 0), "nearest"),
]


@pytest.mark.parametrize("case", OP_CASES, ids=[f"{c[0]}-hd{c[2]}-{c[3]}-{c[4]}-{c[5]}-{c[6]}" for c in OP_CASES])
def test_cuboid_attention_op_vs_oracle(case):
    dims, heads, hd, size, strat, shift, pad = case
    C, B = heads * hd, 2
    g = torch.Generator().manual_seed(7)
    qkv = torch.randn(B, *dims, 3 * C, generator=g).bfloat16()
    n_rel = (2 * size[0] - 1) * (2 * size[1] - 1) * (2 * size[2] - 1)
    table = 0.5 * torch.randn(n_rel, heads, generator=g)
    out = run_op(qkv.cuda(), table.cuda(), dims, C, heads, size, strat, shift, pad)
    ref = O.cuboid_attention_core(qkv.float(), table, heads, size, tuple(strat), shift, pad)
    assert torch.isfinite(out.float()).all()   # every real token written exactly once (output pre-filled with NaN)
    rel_rms, mx = errs(out.float(), ref)
    assert rel_rms < 6e-3 and mx < 1e-2, (rel_rms, mx)


@pytest.mark.parametrize("axis,T,H,W,C", [(0, 13, 16, 16, 256), (1, 13, 16, 16, 256), (2, 13, 8, 8, 512)])
def test_general_kernel_equals_axial_fast_path(axis, T, H, W, C):
    heads, B = 4, 2
    g = torch.Generator().manual_seed(9)
    qkv = torch.randn(B, T, H, W, 3 * C, generator=g).bfloat16().cuda()
    Lx = (T, H, W)[axis]
    table = (0.5 * torch.randn(2 * Lx - 1, heads, generator=g)).cuda()
    size = [1, 1, 1]
    size[axis] = Lx
    a = run_op(qkv, table, (T, H, W), C, heads, size, "lll", (0, 0, 0), "zeros")
    b = torch.empty_like(a)
    L.check(L.lib().pd_op_axial_attention(L.ptr(qkv), L.ptr(table), L.ptr(b), B, T, H, W, C, heads, axis, L.stream_ptr()))
    torch.cuda.synchronize()
    rel_rms, mx = errs(a.float(), b.float().cpu())
    assert rel_rms < 6e-3 and mx < 1.5e-2, (rel_rms, mx)   # two bf16 roundings of the same fp32 math


def make_unet(cfg, max_batch=2):
    m = CuboidTransformerUNet(input_shape=[cfg.t_in, cfg.h, cfg.w, cfg.c], target_shape=[cfg.t_out, cfg.h, cfg.w, cfg.c],
                              base_units=cfg.base_units, depth=list(cfg.depth), num_heads=cfg.num_heads,
                              block_attn_patterns=list(cfg.patterns), padding_type=cfg.padding_type, max_batch=max_batch)
    sd = O.to_torch_sd(Wt.seeded_state_dict(Wt.unet_param_spec(cfg), UNET_SEED))
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all(k.endswith("relative_position_index") for k in res.missing_keys)
    return m.eval(), sd


@pytest.mark.parametrize("case", PC.UNET_CASES + PC.NEAREST_UNET_CASES, ids=[c[0] for c in PC.UNET_CASES + PC.NEAREST_UNET_CASES])
def test_unet_patterns_vs_reference_golden(case):
    tag, pats, pad = case
    cfg = dataclasses.replace(Wt.TINY_UNET, patterns=tuple(pats), padding_type=pad)
    m, _ = make_unet(cfg)
    x = inp(1234, 1, cfg.t_out, cfg.h, cfg.w, cfg.c).cuda()
    cond = inp(1235, 1, cfg.t_in, cfg.h, cfg.w, cfg.c).cuda()
    out = m(x, torch.tensor([500], device="cuda"), cond)
    rel_rms, mx = errs(out, (GN if pad == "nearest" else G)[f"unet_{tag}"])
    assert rel_rms < REL_RMS_TOL and mx < MAX_TOL, (rel_rms, mx)


@pytest.mark.parametrize("pats,pad", [(("video_swin_2x8", "video_swin_2x8"), "ignore"),
                                      (("video_swin_2x8", "video_swin_2x8"), "nearest"),
                                      (("divided_st", "axial_space_dilate_2"), "zeros"),
                                      (("full", "spatial_lg_4"), "zeros")],
                         ids=["swin2x8", "swin2x8_nearest", "dst_dilate", "full_lg"])
def test_unet_shipped_width_patterns_vs_oracle(pats, pad):
    """Width 256 / 512 (head dims 64 / 128; fused projection+FFN kernel at level 0, cluster LayerNorm at level 1) with
    1, 2, 3 and 5 attention layers per stack block."""
    cfg = Wt.UNetConfig(depth=(1, 1), patterns=tuple(pats), padding_type=pad)
    m, sd = make_unet(cfg)
    B = 2
    x = inp(31, B, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cond = inp(32, B, cfg.t_in, cfg.h, cfg.w, cfg.c)
    t = torch.tensor([981, 3])
    out = m(x.cuda(), t.cuda(), cond.cuda())
    with torch.no_grad():
        ref = O.unet_forward(sd, cfg, x, t, cond)
    rel_rms, mx = errs(out, ref)
    assert rel_rms < REL_RMS_TOL and mx < MAX_TOL, (rel_rms, mx)
    # a batch and its shards give the same rows (no cross-sample op)
    out1 = m(x[1:].cuda(), t[1:].cuda(), cond[1:].cuda())
    assert torch.equal(out1[0], out[1])


def test_unet_explicit_cuboid_lists_vs_oracle():
    """block_attn_patterns=None with the reference's default block_cuboid_size / strategy / shift_size
    (cuboid_transformer_unet.py:35-40: (4,4,4) 'l' then (4,4,4) 'd'; T = 13 is padded to 16 under the dilated split)."""
    base = Wt.TINY_UNET
    m = CuboidTransformerUNet(input_shape=[base.t_in, base.h, base.w, base.c], target_shape=[base.t_out, base.h, base.w, base.c],
                              base_units=base.base_units, depth=list(base.depth), num_heads=base.num_heads,
                              block_attn_patterns=None, padding_type="ignore", max_batch=2)
    cfg = m.cfg
    assert cfg.layers(0) == [((4, 4, 4), ("l", "l", "l"), (0, 0, 0)), ((4, 4, 4), ("d", "d", "d"), (0, 0, 0))]
    sd = O.to_torch_sd(Wt.seeded_state_dict(Wt.unet_param_spec(cfg), UNET_SEED))
    m.load_state_dict(sd, strict=False)
    x = inp(41, 2, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cond = inp(42, 2, cfg.t_in, cfg.h, cfg.w, cfg.c)
    t = torch.tensor([700, 50])
    out = m.eval()(x.cuda(), t.cuda(), cond.cuda())
    with torch.no_grad():
        ref = O.unet_forward(sd, cfg, x, t, cond)
    rel_rms, mx = errs(out, ref)
    assert rel_rms < REL_RMS_TOL and mx < MAX_TOL, (rel_rms, mx)
