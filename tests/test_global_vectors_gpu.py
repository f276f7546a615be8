"""GPU parity of the global vectors (cuboid_transformer.py:864-945; cuboid_transformer_unet.py:124-126, 432-434, 449-450,
489-490), through the C ABI:
  - pd_op_gv_linear (the fp32 linears of the K global rows: LayerNorm prologue, GELU, residual) vs torch fp32;
  - pd_op_cuboid_attention_gv vs the reference-pinned oracle core on the same q|k|v: the token grid's output (every query
    also sees the K global keys) and the new global vectors (global queries over all slots, 'ignore' slot mask, self-attn);
  - CuboidTransformerUNet(num_global_vectors > 0) vs goldens of the unmodified reference UNet (tiny config) and vs the
    oracle at the shipped widths; shard invariance; the device-resident DDIM loop (CUDA graph) vs the oracle loop.
Tolerances: fp32 linears and the global queries' attention (fp32 scores, probabilities and accumulation) 2e-5; the token
grid's attention with bf16 probabilities: rel-RMS 6e-3, max 1e-2 of the abs-max; UNet forward rel-RMS 1.5e-2 / max 4e-2 (the
bars of tests/test_patterns_gpu.py; measured 5.7e-3 .. 7.3e-3)."""
import ctypes
import dataclasses
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import prediff_oracle as O
from prediff_b200 import _lib as L
from prediff_b200 import weights as Wt
from prediff_b200.diffusion import LatentDiffusion
from prediff_b200.unet import CuboidTransformerUNet
from tests.golden import pattern_cases as PC
from tests.golden.gen_golden import UNET_SEED, inp

pytestmark = pytest.mark.gpu
GV = np.load(os.path.join(os.path.dirname(__file__), "golden", "global_vectors.npz"))
REL_RMS_TOL, MAX_TOL = 1.5e-2, 4e-2
I3 = ctypes.c_int32 * 3


def errs(out, ref):
    out = out.detach().double().cpu()
    ref = torch.as_tensor(np.asarray(ref)).double()
    rel_rms = ((out - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    mx = ((out - ref).abs().max() / ref.abs().max()).item()
    return rel_rms, mx


@pytest.mark.parametrize("M,K,N,ln,act,res", [(8, 64, 192, True, 0, False), (32, 256, 768, True, 0, False),
                                              (32, 256, 256, False, 0, True), (32, 256, 1024, True, 1, False),
                                              (32, 1024, 256, False, 0, True), (19, 512, 2048, True, 1, False),
                                              (19, 2048, 512, False, 0, True), (130, 128, 67, False, 1, True)])
def test_gv_linear_vs_torch(M, K, N, ln, act, res):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = 0.1 * torch.randn(N, generator=g)
    gamma, beta = 1 + 0.1 * torch.randn(K, generator=g), 0.1 * torch.randn(K, generator=g)
    r = torch.randn(M, N, generator=g)
    y = F.layer_norm(x, (K,), gamma, beta, 1e-5) if ln else x
    want = F.linear(y.double(), W.double(), b.double())
    if act:
        want = F.gelu(want)
    if res:
        want = want + r.double()
    out = r.clone().cuda() if res else torch.full((M, N), float("nan"), device="cuda")
    out16 = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    xd, gd, bd, Wd, biasd = x.cuda(), gamma.cuda(), beta.cuda(), W.cuda(), b.cuda()   # kept alive across the call
    L.check(L.lib().pd_op_gv_linear(L.ptr(xd), L.ptr(gd) if ln else None, L.ptr(bd) if ln else None, L.ptr(Wd), L.ptr(biasd),
                                    L.ptr(out) if res else None, L.ptr(out), L.ptr(out16), M, K, N, act, L.stream_ptr()))
    torch.cuda.synchronize()
    rel_rms, mx = errs(out, want)
    assert rel_rms < 2e-5 and mx < 2e-5, (rel_rms, mx)
    assert torch.equal(out16, out.bfloat16())


# (dims, heads, head_dim, cuboid, strategy, shift, padding, K, global self-attention)
OP_CASES = [
    ((13, 16, 16), 4, 64, (13, 1, 1), "lll", (0, 0, 0), "zeros", 8, False),       # axial layer of the shipped config + globals
    ((13, 8, 8), 4, 128, (1, 8, 1), "lll", (0, 0, 0), "zeros", 8, True),
    ((13, 16, 16), 4, 64, (4, 4, 4), "lll", (2, 2, 2), "zeros", 4, False),        # padded slots (zero rows) take part
    ((13, 16, 16), 4, 64, (4, 4, 4), "lll", (2, 2, 2), "ignore", 4, True),        # raster-order slot mask, shifted
    ((13, 8, 8), 4, 128, (2, 8, 8), "lll", (1, 4, 4), "ignore", 16, True),        # volume 128: two key chunks + the global chunk
    ((13, 8, 8), 4, 32, (2, 4, 4), "ddd", (0, 0, 0), "ignore", 32, False),        # dilated: mask order != slot order; K = 32
    ((6, 7, 9), 2, 16, (4, 3, 4), "ldl", (2, 1, 2), "zeros", 3, True),            # ragged everything
    ((6, 7, 9), 2, 16, (4, 3, 4), "ldl", (2, 1, 2), "ignore", 5, False),
    ((13, 16, 16), 4, 64, (4, 4, 4), "lll", (2, 2, 2), "nearest", 4, True),       # resampled copies take part
    ((13, 16, 16), 4, 64, (1, 16, 16), "lll", (0, 0, 0), "zeros", 8, True),       # divided_st plane: 4 chunks + globals
    ((13, 16, 16), 4, 64, (13, 16, 16), "lll", (0, 0, 0), "ignore", 8, True),     # full: one 3328-slot cuboid, 13 splits
]


@pytest.mark.parametrize("case", OP_CASES, ids=[f"{c[0]}-hd{c[2]}-{c[3]}-{c[4]}-{c[5]}-{c[6]}-K{c[7]}-{int(c[8])}" for c in OP_CASES])
def test_cuboid_attention_gv_op_vs_oracle(case):
    dims, heads, hd, size, strat, shift, pad, K, gsa = case
    C, B = heads * hd, 2
    g = torch.Generator().manual_seed(17)
    qkv = torch.randn(B, *dims, 3 * C, generator=g).bfloat16()
    gqkv = torch.randn(B, K, 3 * C, generator=g).bfloat16().float()   # bf16-representable: both copies carry the same values
    n_rel = (2 * size[0] - 1) * (2 * size[1] - 1) * (2 * size[2] - 1)
    table = 0.5 * torch.randn(n_rel, heads, generator=g)
    out = torch.full((B, *dims, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    gout = torch.full((B, K, C), float("nan"), device="cuda")

    qd, td, gd, gd16 = qkv.cuda(), table.cuda(), gqkv.cuda(), gqkv.bfloat16().cuda()   # kept alive across the calls

    def run():
        L.check(L.lib().pd_op_cuboid_attention_gv(L.ptr(qd), L.ptr(td), L.ptr(gd), L.ptr(gd16), L.ptr(out), L.ptr(gout), B, *dims, C, heads,
                                                  I3(*size), I3(*[0 if s == "l" else 1 for s in strat]), I3(*shift),
                                                  {"zeros": 0, "ignore": 1, "nearest": 2}[pad], K, int(gsa), L.stream_ptr()))
        torch.cuda.synchronize()
    run()
    ref, gref = O.cuboid_attention_core(qkv.float(), table, heads, size, tuple(strat), shift, pad, gqkv=gqkv, global_self_attn=gsa)
    assert torch.isfinite(out.float()).all() and torch.isfinite(gout).all()
    r, m = errs(out.float(), ref)
    rg, mg = errs(gout, gref)
    print(f"cuboid attention + globals {case}: grid rel_rms={r:.2e} max={m:.2e}; global vectors rel_rms={rg:.2e} max={mg:.2e}")
    assert r < 6e-3 and m < 1e-2, (r, m)
    assert rg < 2e-5 and mg < 2e-5, (rg, mg)   # fp32 scores / probabilities / accumulation (measured: 5e-7 / 1.6e-6)
    o1, g1 = out.clone(), gout.clone()
    run()
    assert torch.equal(o1, out) and torch.equal(g1, gout)   # deterministic (fixed split order)


@pytest.mark.parametrize("axis,T,H,W,C,K", [(0, 13, 16, 16, 256, 8), (1, 13, 16, 16, 256, 16), (2, 13, 8, 8, 512, 8),
                                            (0, 13, 8, 8, 512, 3), (2, 6, 8, 12, 64, 1)])
def test_axial_attention_with_global_keys_vs_oracle_and_general_kernel(axis, T, H, W, C, K):
    """The axial fast-path kernel with the global keys as a second key tile: vs the oracle core and vs the general kernel
    (same fp32 math, two bf16 roundings)."""
    heads, B = 4, 2
    g = torch.Generator().manual_seed(23)
    qkv = torch.randn(B, T, H, W, 3 * C, generator=g).bfloat16()
    gqkv = torch.randn(B, K, 3 * C, generator=g).bfloat16()
    Lx = (T, H, W)[axis]
    table = 0.5 * torch.randn(2 * Lx - 1, heads, generator=g)
    size = [1, 1, 1]
    size[axis] = Lx
    qd, td, gd = qkv.cuda(), table.cuda(), gqkv.cuda()
    out = torch.full((B, T, H, W, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    L.check(L.lib().pd_op_axial_attention_gv(L.ptr(qd), L.ptr(td), L.ptr(gd), L.ptr(out), B, T, H, W, C, heads, axis, K,
                                             L.stream_ptr()))
    torch.cuda.synchronize()
    ref, _ = O.cuboid_attention_core(qkv.float(), table, heads, size, ("l", "l", "l"), (0, 0, 0), "zeros", gqkv=gqkv.float())
    assert torch.isfinite(out.float()).all()
    r, m = errs(out.float(), ref)
    gen = torch.empty_like(out)
    gout = torch.empty(B, K, C, device="cuda")
    gd32 = gqkv.float().cuda()
    L.check(L.lib().pd_op_cuboid_attention_gv(L.ptr(qd), L.ptr(td), L.ptr(gd32), L.ptr(gd), L.ptr(gen), L.ptr(gout), B, T, H, W, C,
                                              heads, I3(*size), I3(0, 0, 0), I3(0, 0, 0), 0, K, 0, L.stream_ptr()))
    torch.cuda.synchronize()
    r2, m2 = errs(out.float(), gen.float().cpu())
    print(f"axial + {K} global keys, axis {axis}: vs oracle rel_rms={r:.2e} max={m:.2e}; vs general kernel {r2:.2e} / {m2:.2e}")
    assert r < 6e-3 and m < 1e-2, (r, m)
    assert r2 < 6e-3 and m2 < 1.5e-2, (r2, m2)


# separate_global_qkv=True: (dims, heads, head_dim, cuboid, strategy, shift, padding, K, self-attention, line kernel)
SEP_CASES = [
    ((13, 16, 16), 4, 64, (13, 1, 1), "lll", (0, 0, 0), "zeros", 8, True, True),      # shipped flags on an axial layer
    ((13, 8, 8), 4, 128, (1, 1, 8), "lll", (0, 0, 0), "zeros", 16, False, True),
    ((13, 16, 16), 4, 64, (13, 1, 1), "lll", (0, 0, 0), "zeros", 8, True, False),     # the same layer through the general kernel
    ((13, 16, 16), 4, 64, (4, 4, 4), "lll", (2, 2, 2), "ignore", 4, True, False),
    ((13, 8, 8), 4, 128, (2, 8, 8), "lll", (1, 4, 4), "ignore", 32, True, False),     # two key chunks + 32 global keys
    ((6, 7, 9), 2, 16, (4, 3, 4), "ldl", (2, 1, 2), "nearest", 3, False, False),
    ((13, 8, 8), 4, 32, (2, 4, 4), "ddd", (0, 0, 0), "zeros", 5, True, False),
]


@pytest.mark.parametrize("case", SEP_CASES, ids=[f"{c[0]}-hd{c[2]}-{c[3]}-{c[4]}-{c[5]}-{c[6]}-K{c[7]}-{int(c[8])}-{int(c[9])}" for c in SEP_CASES])
def test_separate_global_qkv_op_vs_oracle(case):
    dims, heads, hd, size, strat, shift, pad, K, gsa, line = case
    C, B = heads * hd, 2
    g = torch.Generator().manual_seed(29)
    qkv = torch.randn(B, *dims, 3 * C, generator=g).bfloat16()
    tok2 = torch.randn(B, *dims, 3 * C, generator=g).bfloat16()                    # l2g_q | g2l_k | g2l_v
    ld = (6 if gsa else 3) * C
    grow = torch.randn(B, K, ld, generator=g).bfloat16().float()                   # l2g_k | l2g_v | g2l_q [| g2g_q | g2g_k | g2g_v]
    n_rel = (2 * size[0] - 1) * (2 * size[1] - 1) * (2 * size[2] - 1)
    table = 0.5 * torch.randn(n_rel, heads, generator=g)
    out = torch.full((B, *dims, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    gout = torch.full((B, K, C), float("nan"), device="cuda")
    qd, t2d, td, gd, gd16 = qkv.cuda(), tok2.cuda(), table.cuda(), grow.cuda(), grow.bfloat16().cuda()
    L.check(L.lib().pd_op_cuboid_attention_gv2(L.ptr(qd), L.ptr(td), L.ptr(t2d), L.ptr(gd), L.ptr(gd16), ld, L.ptr(out), L.ptr(gout),
                                               B, *dims, C, heads, I3(*size), I3(*[0 if s == "l" else 1 for s in strat]), I3(*shift),
                                               {"zeros": 0, "ignore": 1, "nearest": 2}[pad], K, int(gsa), int(line), L.stream_ptr()))
    torch.cuda.synchronize()
    ref, gref = O.cuboid_attention_core(qkv.float(), table, heads, size, tuple(strat), shift, pad, global_self_attn=gsa,
                                        sep=(tok2.float(), grow))
    assert torch.isfinite(out.float()).all() and torch.isfinite(gout).all()
    r, m = errs(out.float(), ref)
    rg, mg = errs(gout, gref)
    print(f"separate global nets {case}: grid rel_rms={r:.2e} max={m:.2e}; global vectors rel_rms={rg:.2e} max={mg:.2e}")
    assert r < 6e-3 and m < 1e-2, (r, m)
    assert rg < 2e-5 and mg < 2e-5, (rg, mg)


GVS = np.load(os.path.join(os.path.dirname(__file__), "golden", "global_vectors_sep.npz"))


@pytest.mark.parametrize("case", PC.GV_SEP_UNET_CASES, ids=[c[0] for c in PC.GV_SEP_UNET_CASES])
def test_unet_separate_global_qkv_vs_reference_golden(case):
    cfg = dataclasses.replace(gv_cfg(case), separate_global_qkv=True)
    m, _ = make_unet(cfg)
    x = inp(1234, 2, cfg.t_out, cfg.h, cfg.w, cfg.c).cuda()
    cond = inp(1235, 2, cfg.t_in, cfg.h, cfg.w, cfg.c).cuda()
    out = m(x, torch.tensor([500, 37], device="cuda"), cond)
    rel_rms, mx = errs(out, GVS[f"unet_{case[0]}"])
    print(f"unet {case[0]}: rel_rms={rel_rms:.2e} max={mx:.2e}")
    assert rel_rms < REL_RMS_TOL and mx < MAX_TOL, (rel_rms, mx)


def test_unet_shipped_flags_with_global_vectors_vs_oracle():
    """The shipped cfg.yaml's global-vector flags (separate_global_qkv: true, use_global_self_attn: true,
    use_global_vector_ffn: false) with num_global_vectors turned on, at the shipped widths."""
    cfg = dataclasses.replace(Wt.UNetConfig(depth=(1, 1)), num_global_vectors=8, use_global_vector_ffn=False,
                              use_global_self_attn=True, separate_global_qkv=True)
    m, sd = make_unet(cfg)
    x, cond, t = inp(31, 2, cfg.t_out, cfg.h, cfg.w, cfg.c), inp(32, 2, cfg.t_in, cfg.h, cfg.w, cfg.c), torch.tensor([981, 3])
    out = m(x.cuda(), t.cuda(), cond.cuda())
    with torch.no_grad():
        ref = O.unet_forward(sd, cfg, x, t, cond)
    rel_rms, mx = errs(out, ref)
    print(f"unet shipped global flags (256 / 512): rel_rms={rel_rms:.2e} max={mx:.2e}")
    assert rel_rms < REL_RMS_TOL and mx < MAX_TOL, (rel_rms, mx)
    assert torch.equal(m(x[1:].cuda(), t[1:].cuda(), cond[1:].cuda())[0], out[1])


def make_unet(cfg, max_batch=2):
    m = CuboidTransformerUNet(input_shape=[cfg.t_in, cfg.h, cfg.w, cfg.c], target_shape=[cfg.t_out, cfg.h, cfg.w, cfg.c],
                              base_units=cfg.base_units, depth=list(cfg.depth), num_heads=cfg.num_heads,
                              block_attn_patterns=list(cfg.patterns), padding_type=cfg.padding_type, max_batch=max_batch,
                              num_global_vectors=cfg.num_global_vectors, use_global_vector_ffn=cfg.use_global_vector_ffn,
                              use_global_self_attn=cfg.use_global_self_attn, separate_global_qkv=cfg.separate_global_qkv)
    sd = O.to_torch_sd(Wt.seeded_state_dict(Wt.unet_param_spec(cfg), UNET_SEED))
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all(k.endswith("relative_position_index") for k in res.missing_keys)
    return m.eval(), sd


def gv_cfg(case, base=Wt.TINY_UNET):
    tag, pats, pad, K, gffn, gsa = case
    return dataclasses.replace(base, patterns=tuple(pats), padding_type=pad, num_global_vectors=K, use_global_vector_ffn=gffn,
                               use_global_self_attn=gsa)


@pytest.mark.parametrize("case", PC.GV_UNET_CASES, ids=[c[0] for c in PC.GV_UNET_CASES])
def test_unet_global_vectors_vs_reference_golden(case):
    cfg = gv_cfg(case)
    m, _ = make_unet(cfg)
    x = inp(1234, 2, cfg.t_out, cfg.h, cfg.w, cfg.c).cuda()
    cond = inp(1235, 2, cfg.t_in, cfg.h, cfg.w, cfg.c).cuda()
    out = m(x, torch.tensor([500, 37], device="cuda"), cond)
    rel_rms, mx = errs(out, GV[f"unet_{case[0]}"])
    print(f"unet {case[0]}: rel_rms={rel_rms:.2e} max={mx:.2e}")
    assert rel_rms < REL_RMS_TOL and mx < MAX_TOL, (rel_rms, mx)


@pytest.mark.parametrize("case", [("w_axial", ("axial", "axial"), "zeros", 8, True, False),
                                  ("w_axial24", ("axial", "axial"), "zeros", 24, False, True),   # > 16: the general kernel
                                  ("w_swin", ("video_swin_2x8", "spatial_lg_4"), "ignore", 8, True, True)],
                         ids=["axial", "axial_k24", "swin2x8_lg"])
def test_unet_shipped_width_global_vectors_vs_oracle(case):
    """Widths 256 / 512 (head dims 64 / 128; the fused projection + FFN kernels keep running beside the global path)."""
    cfg = gv_cfg(case, Wt.UNetConfig(depth=(1, 1)))
    m, sd = make_unet(cfg)
    B = 2
    x = inp(31, B, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cond = inp(32, B, cfg.t_in, cfg.h, cfg.w, cfg.c)
    t = torch.tensor([981, 3])
    out = m(x.cuda(), t.cuda(), cond.cuda())
    with torch.no_grad():
        ref = O.unet_forward(sd, cfg, x, t, cond)
    rel_rms, mx = errs(out, ref)
    print(f"unet {case[0]} (256 / 512): rel_rms={rel_rms:.2e} max={mx:.2e}")
    assert rel_rms < REL_RMS_TOL and mx < MAX_TOL, (rel_rms, mx)
    out1 = m(x[1:].cuda(), t[1:].cuda(), cond[1:].cuda())   # a batch and its shards give the same rows
    assert torch.equal(out1[0], out[1])


def test_ddim_loop_with_global_vectors_vs_oracle():
    """The device-resident loop (one CUDA graph per loop) with the global path's kernels inside, vs the oracle loop."""
    cfg = gv_cfg(PC.GV_UNET_CASES[0])
    m, sd = make_unet(cfg)
    ldm = LatentDiffusion(torch_nn_module=m)
    z = inp(4242, 2, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cond = inp(4243, 2, cfg.t_in, cfg.h, cfg.w, cfg.c)
    z0 = ldm.ddim_sample_loop(cond=cond.cuda(), shape=tuple(z.shape), x_T=z.cuda(), ddim_steps=5, eta=0.0)
    with torch.no_grad():
        ref = O.sample_loop_ddim(sd, cfg, O.make_schedule(), z.clone(), cond, 5)
    r, mx = errs(z0, ref)
    print(f"ddim 5 steps with global vectors: rel_rms={r:.2e} max={mx:.2e}")
    assert r < 1.2e-2 and mx < 2.5e-2
    assert torch.equal(z0, ldm.ddim_sample_loop(cond=cond.cuda(), shape=tuple(z.shape), x_T=z.cuda(), ddim_steps=5, eta=0.0))


def test_side_lane_changes_nothing_and_is_repeatable():
    """The global rows' kernels run on a second plan lane (Plan::lane / mark / wait). Its result must equal the single-stream
    plan's bit for bit (computed by a child process with PD_NO_GV_LANES=1: the switch is read once per process) and must not
    vary between replays (a missing dependency between the lanes would show up as run-to-run differences)."""
    import subprocess
    import sys
    import tempfile
    cfg = dataclasses.replace(Wt.UNetConfig(depth=(1, 1)), num_global_vectors=8, use_global_vector_ffn=True,
                              use_global_self_attn=True, separate_global_qkv=True)
    m, _ = make_unet(cfg)
    x, cond, t = inp(51, 2, cfg.t_out, cfg.h, cfg.w, cfg.c).cuda(), inp(52, 2, cfg.t_in, cfg.h, cfg.w, cfg.c).cuda(), torch.tensor([700, 11]).cuda()
    first = m(x, t, cond)
    for _ in range(30):
        assert torch.equal(m(x, t, cond), first)
    ldm = LatentDiffusion(torch_nn_module=m)
    z0 = ldm.ddim_sample_loop(cond=cond, shape=tuple(x.shape), x_T=x, ddim_steps=5, eta=0.0)   # lanes as graph branches
    for _ in range(5):
        assert torch.equal(ldm.ddim_sample_loop(cond=cond, shape=tuple(x.shape), x_T=x, ddim_steps=5, eta=0.0), z0)
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "single_lane.pt")
        code = (
            "import dataclasses, torch\n"
            "from prediff_b200 import weights as Wt\n"
            "from tests.test_global_vectors_gpu import make_unet, inp\n"
            "cfg = dataclasses.replace(Wt.UNetConfig(depth=(1, 1)), num_global_vectors=8, use_global_vector_ffn=True,\n"
            "                          use_global_self_attn=True, separate_global_qkv=True)\n"
            "m, _ = make_unet(cfg)\n"
            "x, cond = inp(51, 2, cfg.t_out, cfg.h, cfg.w, cfg.c).cuda(), inp(52, 2, cfg.t_in, cfg.h, cfg.w, cfg.c).cuda()\n"
            f"torch.save(m(x, torch.tensor([700, 11]).cuda(), cond).cpu(), {out!r})\n")
        env = dict(os.environ, PD_NO_GV_LANES="1")
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd=root, timeout=600)
        assert torch.equal(torch.load(out), first.cpu())
