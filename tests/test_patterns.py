"""CPU tests for the non-axial cuboid patterns: pins the oracle's general cuboid attention and
the pattern-configured UNet against outputs of the unmodified reference (tests/golden/patterns.npz), and checks the
host-side geometry tables the CUDA kernel consumes (prediff_b200/patterns.py) against the pinned oracle."""
import dataclasses
import os

import numpy as np
import pytest
import torch

from oracle import prediff_oracle as O
from prediff_b200 import patterns as P
from prediff_b200 import weights as Wt
from tests.golden import pattern_cases as PC
from tests.golden.gen_golden import UNET_SEED, inp

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "patterns.npz"))
GN = np.load(os.path.join(os.path.dirname(__file__), "golden", "patterns_nearest.npz"))   # padding_type='nearest'


def maxrel(a, b):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    return ((a - b).abs().max() / b.abs().max()).item()


def test_pattern_registry_names():
    # names registered by cuboid_transformer_patterns.py:60-118
    for name in ("full", "axial", "video_swin", "divided_st", "spatial_lg_v1", "video_swin_2x8", "video_swin_10x32",
                 "spatial_lg_4", "axial_space_dilate_4"):
        assert P.resolve(name, (13, 16, 16, 256))
    for bad in ("video_swin_3x4", "spatial_lg_5", "axial_space_dilate_3", "nope"):
        with pytest.raises(KeyError):
            P.get(bad)
    assert P.resolve("video_swin_2x8", (13, 8, 8, 512)) == [((2, 8, 8), ("l", "l", "l"), (0, 0, 0)),
                                                             ((2, 8, 8), ("l", "l", "l"), (1, 4, 4))]
    assert P.resolve("spatial_lg_8", (13, 8, 8, 512)) == [((13, 1, 1), ("l",) * 3, (0, 0, 0)), ((1, 8, 8), ("l",) * 3, (0, 0, 0))]
    assert [s for s, _, _ in P.resolve("axial_space_dilate_2", (13, 8, 8, 512))] == \
        [(13, 1, 1), (1, 4, 1), (1, 4, 1), (1, 1, 4), (1, 1, 4)]


_ALL_LAYER_CASES = PC.LAYER_CASES + PC.sweep_cases()


@pytest.mark.parametrize("case", _ALL_LAYER_CASES, ids=[c[0] for c in _ALL_LAYER_CASES])
def test_oracle_layer_vs_reference(case):
    tag, dims, C, heads, size, strat, shift, pad = case
    sd = O.to_torch_sd(Wt.seeded_state_dict(PC.layer_spec(C, heads, size), PC.LAYER_SEED))
    x = inp(PC.LAYER_SEED + 1, 2, *dims, C)
    out = O.cuboid_attention(sd, "a", x, heads, size, tuple(strat), shift, pad)
    assert maxrel(out, G[f"layer_{tag}"]) < 2e-5


@pytest.mark.parametrize("case", PC.NEAREST_LAYER_CASES, ids=[c[0] for c in PC.NEAREST_LAYER_CASES])
def test_oracle_layer_nearest_vs_reference(case):
    """padding_type='nearest' (models/utils.py:228-270) against the unmodified CuboidSelfAttentionLayer."""
    tag, dims, C, heads, size, strat, shift, pad = case
    sd = O.to_torch_sd(Wt.seeded_state_dict(PC.layer_spec(C, heads, size), PC.LAYER_SEED))
    x = inp(PC.LAYER_SEED + 1, 2, *dims, C)
    out = O.cuboid_attention(sd, "a", x, heads, size, tuple(strat), shift, pad)
    assert maxrel(out, GN[f"layer_{tag}"]) < 2e-5


_ALL_UNET_CASES = PC.UNET_CASES + [("explicit", None, "ignore")] + PC.NEAREST_UNET_CASES


@pytest.mark.parametrize("case", _ALL_UNET_CASES, ids=[c[0] for c in _ALL_UNET_CASES])
def test_oracle_unet_patterns_vs_reference(case):
    tag, pats, pad = case
    if pats is None:   # block_attn_patterns=None with the reference's default explicit lists
        cfg = dataclasses.replace(Wt.TINY_UNET, patterns=("explicit", "explicit"), padding_type=pad,
                                  explicit_layers=PC.EXPLICIT_LAYERS)
    else:
        cfg = dataclasses.replace(Wt.TINY_UNET, patterns=tuple(pats), padding_type=pad)
    sd = O.to_torch_sd(Wt.seeded_state_dict(Wt.unet_param_spec(cfg), UNET_SEED))
    x = inp(1234, 1, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cond = inp(1235, 1, cfg.t_in, cfg.h, cfg.w, cfg.c)
    out = O.unet_forward(sd, cfg, x, torch.tensor([500]), cond)
    assert maxrel(out, (GN if pad == "nearest" else G)[f"unet_{tag}"]) < 1e-4


GV = np.load(os.path.join(os.path.dirname(__file__), "golden", "global_vectors.npz"))   # use_global_vector=True


@pytest.mark.parametrize("case", PC.GV_LAYER_CASES, ids=[c[0] for c in PC.GV_LAYER_CASES])
def test_oracle_layer_global_vectors_vs_reference(case):
    """Global vectors (cuboid_transformer.py:864-945) against the unmodified CuboidSelfAttentionLayer(use_global_vector=True):
    both outputs, the token grid and the new global vectors."""
    tag, dims, C, heads, size, strat, shift, pad, K, gsa = case
    sd = O.to_torch_sd(Wt.seeded_state_dict(PC.gv_layer_spec(C, heads, size), PC.LAYER_SEED))
    x = inp(PC.LAYER_SEED + 1, 2, *dims, C)
    g = inp(PC.LAYER_SEED + 2, 2, K, C)
    xo, go = O.cuboid_attention_gv(sd, "a", x, g, heads, size, tuple(strat), shift, pad, gsa)
    assert maxrel(xo, GV[f"layer_{tag}_x"]) < 2e-5
    assert maxrel(go, GV[f"layer_{tag}_g"]) < 2e-5


def gv_unet_cfg(case):
    tag, pats, pad, K, gffn, gsa = case
    return dataclasses.replace(Wt.TINY_UNET, patterns=tuple(pats), padding_type=pad, num_global_vectors=K,
                               use_global_vector_ffn=gffn, use_global_self_attn=gsa)


@pytest.mark.parametrize("case", PC.GV_UNET_CASES, ids=[c[0] for c in PC.GV_UNET_CASES])
def test_oracle_unet_global_vectors_vs_reference(case):
    cfg = gv_unet_cfg(case)
    sd = O.to_torch_sd(Wt.seeded_state_dict(Wt.unet_param_spec(cfg), UNET_SEED))
    x = inp(1234, 2, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cond = inp(1235, 2, cfg.t_in, cfg.h, cfg.w, cfg.c)
    out = O.unet_forward(sd, cfg, x, torch.tensor([500, 37]), cond)
    assert maxrel(out, GV[f"unet_{case[0]}"]) < 1e-4


GVS = np.load(os.path.join(os.path.dirname(__file__), "golden", "global_vectors_sep.npz"))   # separate_global_qkv=True


@pytest.mark.parametrize("case", PC.GV_SEP_LAYER_CASES, ids=[c[0] for c in PC.GV_SEP_LAYER_CASES])
def test_oracle_layer_separate_global_qkv_vs_reference(case):
    tag, dims, C, heads, size, strat, shift, pad, K, gsa = case
    sd = O.to_torch_sd(Wt.seeded_state_dict(PC.gv_sep_layer_spec(C, heads, size, gsa), PC.LAYER_SEED))
    x = inp(PC.LAYER_SEED + 1, 2, *dims, C)
    g = inp(PC.LAYER_SEED + 2, 2, K, C)
    xo, go = O.cuboid_attention_gv(sd, "a", x, g, heads, size, tuple(strat), shift, pad, gsa, separate=True)
    assert maxrel(xo, GVS[f"layer_{tag}_x"]) < 2e-5
    assert maxrel(go, GVS[f"layer_{tag}_g"]) < 2e-5


@pytest.mark.parametrize("case", PC.GV_SEP_UNET_CASES, ids=[c[0] for c in PC.GV_SEP_UNET_CASES])
def test_oracle_unet_separate_global_qkv_vs_reference(case):
    cfg = dataclasses.replace(gv_unet_cfg(case), separate_global_qkv=True)
    sd = O.to_torch_sd(Wt.seeded_state_dict(Wt.unet_param_spec(cfg), UNET_SEED))
    x = inp(1234, 2, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cond = inp(1235, 2, cfg.t_in, cfg.h, cfg.w, cfg.c)
    out = O.unet_forward(sd, cfg, x, torch.tensor([500, 37]), cond)
    assert maxrel(out, GVS[f"unet_{case[0]}"]) < 1e-4


def attention_from_tables(qkv, table, heads, geo):
    """What the CUDA kernel computes, in numpy: gather rows by `tok`, mask by `lab`, bias by `rel`."""
    B, T, H, W, C3 = qkv.shape
    C, hd = C3 // 3, C3 // 3 // heads
    nc, vol = geo["num_cuboids"], geo["volume"]
    tok, lab, rel = geo["tok"].reshape(nc, vol), geo["lab"].reshape(nc, vol), geo["rel"]
    rows = np.concatenate([qkv.reshape(B, T * H * W, C3), np.zeros((B, 1, C3), qkv.dtype)], axis=1)
    out = np.zeros((B, T * H * W + 1, C), np.float64)
    bias = table[rel[:, None] - rel[None, :] + geo["rel_off"]]  # (vol, vol, heads)
    for c in range(nc):
        y = rows[:, tok[c]].astype(np.float64).reshape(B, vol, 3, heads, hd)   # tok == -1 picks the zero row
        q, k, v = (y[:, :, i].transpose(0, 2, 1, 3) for i in range(3))
        s = (q * hd ** -0.5) @ k.transpose(0, 1, 3, 2) + bias.transpose(2, 0, 1)[None]
        m = (lab[c][:, None] == lab[c][None, :]) & (lab[c][:, None] >= 0) & (lab[c][None, :] >= 0)
        s = np.where(m, s, -np.inf)
        mx = np.where(np.isfinite(s.max(-1, keepdims=True)), s.max(-1, keepdims=True), 0.0)
        p = np.exp(s - mx)
        den = p.sum(-1, keepdims=True)
        p = np.where(den > 0, p / np.where(den > 0, den, 1.0), 0.0)
        o = (p @ v).transpose(0, 2, 1, 3).reshape(B, vol, C)
        if geo.get("dst") is not None:   # 'nearest': slots write to their destination token (-1 = the scratch row)
            out[:, geo["dst"].reshape(nc, vol)[c]] = o
        else:
            out[:, tok[c]] = o   # padded tokens land in the scratch row
    return out[:, :-1].reshape(B, T, H, W, C)


GEOM_CASES = [(c[1], c[3], c[4], c[5], c[6], c[7]) for c in _ALL_LAYER_CASES + PC.NEAREST_LAYER_CASES] + [
    ((13, 16, 16), 4, (2, 8, 8), "lll", (1, 4, 4), "nearest"),
    ((13, 8, 8), 4, (4, 4, 4), "dld", (0, 2, 0), "nearest"),
    ((13, 16, 16), 4, (13, 1, 1), "lll", (0, 0, 0), "zeros"),
    ((13, 16, 16), 4, (1, 16, 16), "lll", (0, 0, 0), "ignore"),
    ((13, 8, 8), 4, (2, 8, 8), "lll", (1, 4, 4), "ignore"),
    ((13, 8, 8), 4, (1, 4, 1), "ddd", (0, 0, 0), "zeros"),
]


@pytest.mark.parametrize("case", GEOM_CASES, ids=[f"{c[0]}-{c[2]}-{c[3]}-{c[4]}-{c[5]}" for c in GEOM_CASES])
def test_geometry_tables_vs_oracle(case):
    dims, heads, size, strat, shift, pad = case
    hd = 8
    C = heads * hd
    rng = np.random.Generator(np.random.PCG64(11))
    qkv = rng.standard_normal((2, *dims, 3 * C), dtype=np.float32)
    n_rel = (2 * size[0] - 1) * (2 * size[1] - 1) * (2 * size[2] - 1)
    table = (0.3 * rng.standard_normal((n_rel, heads), dtype=np.float32))
    want = O.cuboid_attention_core(torch.from_numpy(qkv), torch.from_numpy(table), heads, size, tuple(strat), shift, pad)
    geo = P.layer_geometry(dims, size, tuple(strat), shift, pad)
    got = attention_from_tables(qkv, table, heads, geo)
    assert maxrel(got, want) < 1e-5
    # every real token belongs to exactly one cuboid slot ('nearest': is written by exactly one slot)
    t = geo["tok"][geo["tok"] >= 0] if geo["dst"] is None else geo["dst"][geo["dst"] >= 0]
    assert np.array_equal(np.sort(t), np.arange(dims[0] * dims[1] * dims[2]))


def global_attention_from_tables(qkv, gqkv, heads, geo, self_attn):
    """What the CUDA path computes for the global vectors, in numpy: every cuboid's queries see the K global keys as extra
    unmasked columns; the global queries attend over the slot list (`tok`, hidden where `gmask` is 0) + themselves."""
    B, T, H, W, C3 = qkv.shape
    C, hd = C3 // 3, C3 // 3 // heads
    K = gqkv.shape[1]
    tok, gm = geo["tok"], geo["gmask"]
    rows = np.concatenate([qkv.reshape(B, T * H * W, C3), np.zeros((B, 1, C3), qkv.dtype)], axis=1).astype(np.float64)
    g = gqkv.astype(np.float64).reshape(B, K, 3, heads, hd)
    y = rows[:, tok].reshape(B, -1, 3, heads, hd)                      # (B, slots, 3, heads, hd), tok == -1: the zero row
    k_all, v_all = y[:, :, 1], y[:, :, 2]
    vis = np.ones(len(tok), bool) if gm is None else gm > 0
    if self_attn:
        k_all, v_all = np.concatenate([k_all, g[:, :, 1]], 1), np.concatenate([v_all, g[:, :, 2]], 1)
        vis = np.concatenate([vis, np.ones(K, bool)])
    s = np.einsum("bqhd,bkhd->bhqk", g[:, :, 0] * hd ** -0.5, k_all)
    s = np.where(vis, s, -np.inf)
    p = np.exp(s - s.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    return np.einsum("bhqk,bkhd->bqhd", p, v_all).reshape(B, K, C)


_GV_GEOM = [(c[1], c[3], c[4], c[5], c[6], c[7], c[8], c[9]) for c in PC.GV_LAYER_CASES]


@pytest.mark.parametrize("case", _GV_GEOM, ids=[c[0] for c in PC.GV_LAYER_CASES])
def test_global_vector_tables_vs_oracle(case):
    """The slot list + gmask the global-attention kernel walks reproduce the pinned oracle's new global vectors."""
    dims, heads, size, strat, shift, pad, K, gsa = case
    hd = 8
    C = heads * hd
    rng = np.random.Generator(np.random.PCG64(13))
    qkv = rng.standard_normal((2, *dims, 3 * C), dtype=np.float32)
    gqkv = rng.standard_normal((2, K, 3 * C), dtype=np.float32)
    n_rel = (2 * size[0] - 1) * (2 * size[1] - 1) * (2 * size[2] - 1)
    table = (0.3 * rng.standard_normal((n_rel, heads), dtype=np.float32))
    _, want = O.cuboid_attention_core(torch.from_numpy(qkv), torch.from_numpy(table), heads, size, tuple(strat), shift, pad,
                                      gqkv=torch.from_numpy(gqkv), global_self_attn=gsa)
    geo = P.layer_geometry(dims, size, tuple(strat), shift, pad)
    assert (geo["gmask"] is not None) == (pad == "ignore")
    got = global_attention_from_tables(qkv, gqkv, heads, geo, gsa)
    assert maxrel(got, want) < 1e-5


@pytest.mark.parametrize("case", GEOM_CASES, ids=[f"{c[0]}-{c[2]}-{c[3]}-{c[4]}-{c[5]}" for c in GEOM_CASES])
def test_c_abi_tables_equal_python_tables(case):
    """pd_cuboid_tables (host-only C++ builder the UNet plan uses) == prediff_b200.patterns.layer_geometry."""
    import ctypes
    from prediff_b200 import _lib as L
    dims, heads, size, strat, shift, pad = case
    geo = P.layer_geometry(dims, size, tuple(strat), shift, pad)
    i3 = ctypes.c_int32 * 3
    meta = (ctypes.c_int32 * 12)()
    n = geo["num_cuboids"] * geo["volume"]
    tok, lab, rel = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(geo["volume"], np.int32)
    rc = L.lib().pd_cuboid_tables(*dims, i3(*size), i3(*[0 if s == "l" else 1 for s in strat]), i3(*shift),
                                  {"zeros": 0, "ignore": 1, "nearest": 2}[pad], meta, tok.ctypes.data_as(ctypes.c_void_p),
                                  lab.ctypes.data_as(ctypes.c_void_p), rel.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(n))
    assert rc >= 0
    assert tuple(meta[0:3]) == geo["size"] and tuple(meta[3:6]) == geo["shift"] and tuple(meta[6:9]) == geo["pad"]
    assert (meta[9], meta[10], meta[11]) == (geo["num_cuboids"], geo["volume"], geo["rel_off"])
    assert np.array_equal(tok, geo["tok"]) and np.array_equal(lab, geo["lab"]) and np.array_equal(rel, geo["rel"])
    dst = np.full(n, -7, np.int32)
    rd = L.lib().pd_cuboid_tables_dst(*dims, i3(*size), i3(*[0 if s == "l" else 1 for s in strat]), i3(*shift),
                                      {"zeros": 0, "ignore": 1, "nearest": 2}[pad], dst.ctypes.data_as(ctypes.c_void_p),
                                      ctypes.c_int64(n))
    assert rd == (0 if geo["dst"] is None else 1)
    if geo["dst"] is not None:
        assert np.array_equal(dst, geo["dst"])
    gm = np.full(n, -7, np.int32)
    rg = L.lib().pd_cuboid_tables_gmask(*dims, i3(*size), i3(*[0 if s == "l" else 1 for s in strat]), i3(*shift),
                                        {"zeros": 0, "ignore": 1, "nearest": 2}[pad], gm.ctypes.data_as(ctypes.c_void_p),
                                        ctypes.c_int64(n))
    assert rg == (0 if geo["gmask"] is None else 1)
    if geo["gmask"] is not None:
        assert np.array_equal(gm, geo["gmask"])
    is_axial = sum(s > 1 for s in geo["size"]) == 1 and geo["size"] == tuple(size) and max(geo["size"]) <= 16 \
        and all(geo["size"][a] in (1, dims[a]) for a in range(3)) and not any(geo["shift"])
    assert (rc > 0) == is_axial
