"""GPU parity: CuboidTransformerUNet CUDA path (through the C ABI) vs golden outputs of the unmodified reference
(tests/golden) and vs the CPU oracle on fresh inputs.

Tolerance: the CUDA path feeds bf16 operands to the tensor cores (fp32 accumulate, fp32 residual stream); the
reference's own GPU numerics are TF32. SURVEY.md section 7 measured bf16-operand emulation at rel-RMS 6.8e-3 per UNet
step; measured on B200 6.4e-3 - 6.9e-3 (profiles/parity_r02.txt). The bars here are ~1.5x the measured values: rel-RMS
<= 1.2e-2 and max-abs error <= 2.5e-2 of the output's abs-max. The like-for-like TF32 mode is held to 2e-3 in test_tf32_gpu.py."""
import os

import numpy as np
import pytest
import torch

from oracle import prediff_oracle as O
from prediff_b200 import weights as Wt
from prediff_b200.unet import CuboidTransformerUNet
from tests.golden.gen_golden import UNET_SEED, inp

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
# measured on B200 (profiles/parity_r02.txt): UNet step rel-RMS 6.4e-3 - 6.9e-3, max 7.7e-3; VAE (same bars) 9e-3 - ~1.5x those
REL_RMS_TOL, MAX_TOL = 1.2e-2, 2.5e-2


def errs(out, ref):
    out = out.detach().double().cpu()
    ref = torch.as_tensor(np.asarray(ref)).double()
    rel_rms = ((out - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    mx = ((out - ref).abs().max() / ref.abs().max()).item()
    return rel_rms, mx


def make_unet(cfg, max_batch=4):
    m = CuboidTransformerUNet(input_shape=[cfg.t_in, cfg.h, cfg.w, cfg.c], target_shape=[cfg.t_out, cfg.h, cfg.w, cfg.c],
                              base_units=cfg.base_units, depth=list(cfg.depth), num_heads=cfg.num_heads,
                              block_attn_patterns="axial", max_batch=max_batch)
    sd = O.to_torch_sd(Wt.seeded_state_dict(Wt.unet_param_spec(cfg), UNET_SEED))
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all(k.endswith("relative_position_index") for k in res.missing_keys)
    return m.eval(), sd


@pytest.fixture(scope="module")
def tiny():
    return make_unet(Wt.TINY_UNET)


def test_unet_tiny_vs_reference_golden(tiny):
    m, _ = tiny
    cfg = Wt.TINY_UNET
    g = np.load(os.path.join(G, "unet_tiny.npz"))
    x = inp(1234, 2, cfg.t_out, cfg.h, cfg.w, cfg.c).cuda()
    cond = inp(1235, 2, cfg.t_in, cfg.h, cfg.w, cfg.c).cuda()
    out = m(x, torch.as_tensor(g["t"]).cuda(), cond)
    rel_rms, mx = errs(out, g["out"])
    print(f"unet tiny vs reference: rel_rms={rel_rms:.3e} max={mx:.3e}")
    assert rel_rms < REL_RMS_TOL and mx < MAX_TOL


@pytest.mark.parametrize("B,ts", [(1, [0]), (3, [999, 1, 250]), (4, [7, 7, 7, 7])])
def test_unet_tiny_vs_oracle_fresh_inputs(tiny, B, ts):
    m, sd = tiny
    cfg = Wt.TINY_UNET
    x = inp(55 + B, B, cfg.t_out, cfg.h, cfg.w, cfg.c)
    cond = inp(66 + B, B, cfg.t_in, cfg.h, cfg.w, cfg.c)
    t = torch.tensor(ts)
    with torch.no_grad():
        ref = O.unet_forward(sd, cfg, x, t, cond)
    out = m(x.cuda(), t.cuda(), cond.cuda())
    rel_rms, mx = errs(out, ref)
    print(f"unet tiny B={B}: rel_rms={rel_rms:.3e} max={mx:.3e}")
    assert rel_rms < REL_RMS_TOL and mx < MAX_TOL


def test_unet_batch_independence(tiny):
    """Samples are independent chains: row b of a batched forward equals the single-sample forward (bit-exact)."""
    m, _ = tiny
    cfg = Wt.TINY_UNET
    x = inp(91, 3, cfg.t_out, cfg.h, cfg.w, cfg.c).cuda()
    cond = inp(92, 3, cfg.t_in, cfg.h, cfg.w, cfg.c).cuda()
    t = torch.tensor([10, 500, 900]).cuda()
    full = m(x, t, cond)
    for b in range(3):
        one = m(x[b:b + 1], t[b:b + 1], cond[b:b + 1])
        assert torch.equal(one[0], full[b])


def test_unet_full_config_vs_reference_golden():
    """BASELINE config 2: single denoise step, batch 1, 13x16x16 latent, shipped cfg.yaml sizes."""
    cfg = Wt.UNetConfig()
    m, _ = make_unet(cfg, max_batch=1)
    g = np.load(os.path.join(G, "unet_full.npz"))
    x = inp(1234, 1, cfg.t_out, cfg.h, cfg.w, cfg.c).cuda()
    cond = inp(1235, 1, cfg.t_in, cfg.h, cfg.w, cfg.c).cuda()
    out = m(x, torch.as_tensor(g["t"]).cuda(), cond)
    rel_rms, mx = errs(out, g["out"])
    print(f"unet full vs reference: rel_rms={rel_rms:.3e} max={mx:.3e}")
    assert rel_rms < REL_RMS_TOL and mx < MAX_TOL


def test_unet_rejects_cpu_tensors(tiny):
    m, _ = tiny
    cfg = Wt.TINY_UNET
    with pytest.raises(Exception):
        m(torch.zeros(1, cfg.t_out, cfg.h, cfg.w, cfg.c), torch.zeros(1, dtype=torch.long),
          torch.zeros(1, cfg.t_in, cfg.h, cfg.w, cfg.c))


def test_streamk_latency_cut_vs_reference_golden():
    """pd_unet_set_streamk_ctas(72): a model built for single samples (more stream-K CTAs per sample; another fixed summation
    order of the convolutions' partial tiles) meets the same parity bar against the unmodified reference's output
    (unet_full.npz, BASELINE config 2) and is deterministic. (The two cuts differ from each other by about as much as either
    differs from the fp32 reference: a 2e-6 difference in an fp32 sum flips bf16 roundings downstream.)"""
    cfg = Wt.UNetConfig()
    sd = O.to_torch_sd(Wt.seeded_state_dict(Wt.unet_param_spec(cfg), UNET_SEED))
    m = CuboidTransformerUNet([cfg.t_in, cfg.h, cfg.w, cfg.c], [cfg.t_out, cfg.h, cfg.w, cfg.c], base_units=cfg.base_units,
                              depth=list(cfg.depth), num_heads=cfg.num_heads, max_batch=1, streamk_ctas_per_sample=72)
    m.load_state_dict(sd, strict=False)
    g = np.load(os.path.join(G, "unet_full.npz"))
    x = inp(1234, 1, cfg.t_out, cfg.h, cfg.w, cfg.c).cuda()
    cond = inp(1235, 1, cfg.t_in, cfg.h, cfg.w, cfg.c).cuda()
    t = torch.as_tensor(g["t"]).cuda()
    out = m(x, t, cond).clone()
    assert torch.equal(out, m(x, t, cond))
    rel_rms, mx = errs(out, g["out"])
    print(f"unet full, latency cut (72 stream-K CTAs per sample) vs reference: rel_rms={rel_rms:.3e} max={mx:.3e}")
    assert rel_rms < REL_RMS_TOL and mx < MAX_TOL
