"""Kernel-level parity tests (GPU): every CUDA kernel, called through the C ABI, against a plain PyTorch fp32
reference of the same op evaluated on the same (bf16-rounded where the kernel consumes bf16) inputs."""
import ctypes
import math

import pytest
import torch
import torch.nn.functional as F

from prediff_b200 import _lib as L

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _init():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    L.init()


def _sync_check(rc):
    L.check(rc)
    torch.cuda.synchronize()


def rel_err(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


def _randn(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,K,N,bn", [
    (128, 64, 64, 0), (256, 256, 256, 0), (3328, 256, 768, 0), (832, 512, 1536, 0), (1000, 128, 96, 0),
    (3328, 1024, 256, 64), (3328, 256, 256, 128), (3328, 256, 256, 256), (512, 2048, 512, 0), (384, 64, 32, 32),
])
def test_gemm_linear(M, K, N, bn):
    a = _randn(M, K, seed=1).bfloat16()
    w = _randn(N, K, seed=2, scale=K ** -0.5).bfloat16()
    bias = _randn(N, seed=3)
    res = _randn(M, N, seed=4)
    # fp32 output with fused bias + GELU + residual (TMA-loaded residual tile, TMA-stored result)
    out = torch.full((M, N), float("nan"), device=DEV)
    _sync_check(L.lib().pd_op_conv_gemm(L.ptr(a), L.ptr(w), 1, 1, 1, M, K, 1, 1, 1, N, L.ptr(bias), None, L.ptr(res),
                                        L.ptr(out), None, 1, bn, L.stream_ptr()))
    ref = F.gelu(a.float() @ w.float().t() + bias) + res
    e = rel_err(out, ref)
    assert e < 2e-5, f"fp32 out rel err {e}"
    # in-place residual (out aliases residual), as the UNet uses it
    out2 = res.clone()
    _sync_check(L.lib().pd_op_conv_gemm(L.ptr(a), L.ptr(w), 1, 1, 1, M, K, 1, 1, 1, N, L.ptr(bias), None, L.ptr(out2),
                                        L.ptr(out2), None, 1, bn, L.stream_ptr()))
    assert torch.equal(out2, out)
    if N % 64 == 0:  # bf16 output path
        outb = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
        _sync_check(L.lib().pd_op_conv_gemm(L.ptr(a), L.ptr(w), 1, 1, 1, M, K, 1, 1, 1, N, L.ptr(bias), None, None,
                                            None, L.ptr(outb), 1, bn, L.stream_ptr()))
        assert rel_err(outb, F.gelu(a.float() @ w.float().t() + bias)) < 1e-2


@pytest.mark.parametrize("M,K,N,act", [(13312, 256, 768, 0), (13312, 256, 1024, 1), (6656, 512, 2048, 1),
                                         (6600, 256, 1024, 1)])
def test_gemm_persistent_bf16(M, K, N, act):
    """Multi-wave bf16-output GEMMs (> 148 tiles of 128 x 256) run the persistent double-buffered-TMEM kernel."""
    a = _randn(M, K, seed=1).bfloat16()
    w = _randn(N, K, seed=2, scale=K ** -0.5).bfloat16()
    bias = _randn(N, seed=3)
    outb = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    _sync_check(L.lib().pd_op_conv_gemm(L.ptr(a), L.ptr(w), 1, 1, 1, M, K, 1, 1, 1, N, L.ptr(bias), None, None, None,
                                        L.ptr(outb), act, 0, L.stream_ptr()))
    ref = a.float() @ w.float().t() + bias
    if act:
        ref = F.gelu(ref)
    assert rel_err(outb, ref) < 1e-2
    # exactness of the accumulation itself: compare against the bf16 rounding of the fp32 reference
    assert (outb.float() - ref.bfloat16().float()).abs().max().item() <= 2 * ref.abs().max().item() * 2 ** -8


@pytest.mark.parametrize("M,K,N", [(3328, 256, 256), (1000, 1024, 256), (13312, 256, 256), (3328, 512, 512),
                                   (832, 2048, 512), (1000, 512, 512)])
def test_gemm_residual_fused_layernorm(M, K, N):
    """N = 512: LayerNorm statistics exchanged between the two CTAs of a cluster through DSMEM."""
    a = _randn(M, K, seed=1).bfloat16()
    w = _randn(N, K, seed=2, scale=K ** -0.5).bfloat16()
    bias = _randn(N, seed=3)
    x = _randn(M, N, seed=4) * 2 + 0.3
    gamma = 1 + 0.1 * _randn(N, seed=5)
    beta = 0.1 * _randn(N, seed=6)
    ref_x = a.float() @ w.float().t() + bias + x
    ref_ln = F.layer_norm(ref_x, (N,), gamma, beta, 1e-5)
    ln = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    _sync_check(L.lib().pd_op_linear_residual_ln_n(L.ptr(a), L.ptr(w), M, K, N, L.ptr(bias), L.ptr(x), L.ptr(gamma),
                                                   L.ptr(beta), L.ptr(ln), L.stream_ptr()))
    assert rel_err(x, ref_x) < 2e-5
    assert rel_err(ln, ref_ln) < 6e-3


@pytest.mark.parametrize("M,with_ln", [(13312, True), (3328, True), (1000, False), (128, True)])
def test_ffn_fused(M, with_ln):
    """One kernel: x += W2 gelu(W1 ln + b1) + b2 (+ LayerNorm of the result); hidden activation rounded to bf16
    between the two GEMMs exactly like the two-kernel path."""
    C, Hd = 256, 1024
    ln_in = _randn(M, C, seed=1).bfloat16()
    w1 = _randn(Hd, C, seed=2, scale=C ** -0.5).bfloat16()
    w2 = _randn(C, Hd, seed=3, scale=Hd ** -0.5).bfloat16()
    b1, b2 = 0.1 * _randn(Hd, seed=4), 0.1 * _randn(C, seed=5)
    x = _randn(M, C, seed=6) * 2 + 0.3
    gamma, beta = 1 + 0.1 * _randn(C, seed=7), 0.1 * _randn(C, seed=8)
    mid = F.gelu(ln_in.float() @ w1.float().t() + b1).bfloat16().float()
    ref_x = x + mid @ w2.float().t() + b2
    ref_ln = F.layer_norm(ref_x, (C,), gamma, beta, 1e-5)
    xo = x.clone()
    ln = torch.zeros(M, C, device=DEV, dtype=torch.bfloat16)
    _sync_check(L.lib().pd_op_ffn_fused(L.ptr(ln_in), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(xo),
                                        L.ptr(gamma) if with_ln else None, L.ptr(beta) if with_ln else None,
                                        L.ptr(ln) if with_ln else None, M, L.stream_ptr()))
    # bf16 rounding of `mid` can flip at ties between the two implementations: tolerance of a few bf16 ulps of mid
    assert rel_err(xo, ref_x) < 2e-3
    if with_ln:
        assert rel_err(ln, ref_ln) < 8e-3
    # determinism
    xo2 = x.clone()
    _sync_check(L.lib().pd_op_ffn_fused(L.ptr(ln_in), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(xo2),
                                        L.ptr(gamma) if with_ln else None, L.ptr(beta) if with_ln else None,
                                        L.ptr(ln) if with_ln else None, M, L.stream_ptr()))
    assert torch.equal(xo, xo2)


@pytest.mark.parametrize("M,rows,with_ln,with_gn", [(3328, 832, True, False), (3328, 832, False, True), (832, 832, True, True),
                                                    (128, 128, True, False), (1664, 832, False, False)])
def test_ffn_cluster(M, rows, with_ln, with_gn):
    """Width-512 FFN in one kernel, hidden dimension split over a 4-CTA cluster with a DSMEM reduce-scatter
    (csrc/ffn_cluster.cu): x += W2 gelu(W1 ln + b1) + b2 (+ LayerNorm of the result, + GroupNorm statistics)."""
    C, Hd = 512, 2048
    ln_in = _randn(M, C, seed=1).bfloat16()
    w1 = _randn(Hd, C, seed=2, scale=C ** -0.5).bfloat16()
    w2 = _randn(C, Hd, seed=3, scale=Hd ** -0.5).bfloat16()
    b1, b2 = 0.1 * _randn(Hd, seed=4), 0.1 * _randn(C, seed=5)
    x = _randn(M, C, seed=6) * 2 + 0.3
    gamma, beta = 1 + 0.1 * _randn(C, seed=7), 0.1 * _randn(C, seed=8)
    mid = F.gelu(ln_in.float() @ w1.float().t() + b1).bfloat16().float()
    ref_x = x + mid @ w2.float().t() + b2
    ref_ln = F.layer_norm(ref_x, (C,), gamma, beta, 1e-5)

    def run():
        xo = x.clone()
        ln = torch.zeros(M, C, device=DEV, dtype=torch.bfloat16)
        sums = torch.zeros(M // rows, 32, 2, device=DEV, dtype=torch.float64)
        _sync_check(L.lib().pd_op_ffn_cluster(L.ptr(ln_in), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(xo),
                                              L.ptr(gamma) if with_ln else None, L.ptr(beta) if with_ln else None,
                                              L.ptr(ln) if with_ln else None, L.ptr(sums) if with_gn else None, 32, rows, M,
                                              L.stream_ptr()))
        return xo, ln, sums

    xo, ln, sums = run()
    assert rel_err(xo, ref_x) < 2e-3     # bf16 rounding of `mid` can flip at ties between the two implementations
    if with_ln:
        assert rel_err(ln, ref_ln) < 8e-3
    if with_gn:
        g = xo.double().reshape(M // rows, rows, 32, C // 32)
        want = torch.stack([g.sum(dim=(1, 3)), (g * g).sum(dim=(1, 3))], dim=-1)
        assert torch.allclose(sums, want, rtol=2e-6, atol=1e-3)
    xo2, ln2, _ = run()                  # deterministic: partials are added in rank order
    assert torch.equal(xo, xo2) and torch.equal(ln, ln2)
    if M > rows:                         # batch invariance: the first sample alone gives the same rows
        M1 = rows
        x1 = x[:M1].clone()
        l1 = torch.zeros(M1, C, device=DEV, dtype=torch.bfloat16)
        a1 = ln_in[:M1].contiguous()
        _sync_check(L.lib().pd_op_ffn_cluster(L.ptr(a1), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(x1),
                                              L.ptr(gamma) if with_ln else None, L.ptr(beta) if with_ln else None,
                                              L.ptr(l1) if with_ln else None, None, 32, rows, M1, L.stream_ptr()))
        assert torch.equal(x1, xo[:M1])


@pytest.mark.parametrize("M,rows,with_ln,with_gn", [(3328, 832, True, False), (3328, 832, False, True), (832, 832, True, True),
                                                    (1000, 1000, True, False)])
def test_proj_ffn_cluster(M, rows, with_ln, with_gn):
    """Width-512 attention projection + residual + pre-norm + FFN (+ next LayerNorm, + GroupNorm statistics) in the cluster
    kernel (csrc/ffn_cluster.cu, PROJ variant): x1 = x + att Wp^T + bp; x = x1 + W2 gelu(W1 LN(x1) + b1) + b2."""
    C, Hd = 512, 2048
    att = _randn(M, C, seed=11).bfloat16()
    wp = _randn(C, C, seed=12, scale=C ** -0.5).bfloat16()
    bp = 0.1 * _randn(C, seed=13)
    g1, be1 = 1 + 0.1 * _randn(C, seed=14), 0.1 * _randn(C, seed=15)
    w1 = _randn(Hd, C, seed=2, scale=C ** -0.5).bfloat16()
    w2 = _randn(C, Hd, seed=3, scale=Hd ** -0.5).bfloat16()
    b1, b2 = 0.1 * _randn(Hd, seed=4), 0.1 * _randn(C, seed=5)
    x = _randn(M, C, seed=6) * 2 + 0.3
    gamma, beta = 1 + 0.1 * _randn(C, seed=7), 0.1 * _randn(C, seed=8)
    x1 = x + att.float() @ wp.float().t() + bp
    ln1 = F.layer_norm(x1, (C,), g1, be1, 1e-5).bfloat16().float()
    mid = F.gelu(ln1 @ w1.float().t() + b1).bfloat16().float()
    ref_x = x1 + mid @ w2.float().t() + b2
    ref_ln = F.layer_norm(ref_x, (C,), gamma, beta, 1e-5)

    def run(Mr=M):
        xo = x[:Mr].clone()
        scratch = torch.zeros(Mr, C, device=DEV, dtype=torch.bfloat16)
        ln = torch.zeros(Mr, C, device=DEV, dtype=torch.bfloat16)
        sums = torch.zeros(max(Mr // rows, 1), 32, 2, device=DEV, dtype=torch.float64)
        a = att[:Mr].contiguous()
        _sync_check(L.lib().pd_op_proj_ffn_cluster(
            L.ptr(a), L.ptr(wp), L.ptr(bp), L.ptr(g1), L.ptr(be1), L.ptr(scratch), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2),
            L.ptr(xo), L.ptr(gamma) if with_ln else None, L.ptr(beta) if with_ln else None, L.ptr(ln) if with_ln else None,
            L.ptr(sums) if with_gn else None, 32, rows, Mr, None, L.stream_ptr()))
        return xo, ln, sums, scratch

    xo, ln, sums, scratch = run()
    assert rel_err(scratch, ln1) < 8e-3  # the pre-norm the kernel published (bf16)
    assert rel_err(xo, ref_x) < 3e-3     # bf16 roundings of LN(x1) / `mid` can flip at ties between the implementations
    if with_ln:
        assert rel_err(ln, ref_ln) < 8e-3
    if with_gn:
        g = xo.double().reshape(M // rows, rows, 32, C // 32)
        want = torch.stack([g.sum(dim=(1, 3)), (g * g).sum(dim=(1, 3))], dim=-1)
        assert torch.allclose(sums, want, rtol=2e-6, atol=1e-3)
    xo2, ln2, _, _ = run()               # deterministic: row sums and partials are combined in rank order
    assert torch.equal(xo, xo2) and torch.equal(ln, ln2)
    if M > rows:                         # batch invariance: the first sample alone gives the same rows
        x1o, _, _, _ = run(rows)
        assert torch.equal(x1o, xo[:rows])


@pytest.mark.parametrize("M,with_ln", [(13312, True), (3328, False), (1000, True)])
def test_proj_ffn_fused(M, with_ln):
    """Attention projection + residual + pre-norm + FFN (+ next LayerNorm) in one kernel."""
    C, Hd = 256, 1024
    att = _randn(M, C, seed=11).bfloat16()
    wp = _randn(C, C, seed=12, scale=C ** -0.5).bfloat16()
    bp = 0.1 * _randn(C, seed=13)
    g1, be1 = 1 + 0.1 * _randn(C, seed=14), 0.1 * _randn(C, seed=15)
    w1 = _randn(Hd, C, seed=2, scale=C ** -0.5).bfloat16()
    w2 = _randn(C, Hd, seed=3, scale=Hd ** -0.5).bfloat16()
    b1, b2 = 0.1 * _randn(Hd, seed=4), 0.1 * _randn(C, seed=5)
    x = _randn(M, C, seed=6) * 2 + 0.3
    gamma, beta = 1 + 0.1 * _randn(C, seed=7), 0.1 * _randn(C, seed=8)
    x1 = x + att.float() @ wp.float().t() + bp
    ln1 = F.layer_norm(x1, (C,), g1, be1, 1e-5).bfloat16().float()
    mid = F.gelu(ln1 @ w1.float().t() + b1).bfloat16().float()
    ref_x = x1 + mid @ w2.float().t() + b2
    ref_ln = F.layer_norm(ref_x, (C,), gamma, beta, 1e-5)
    xo = x.clone()
    ln = torch.zeros(M, C, device=DEV, dtype=torch.bfloat16)
    args = lambda xx: (L.ptr(att), L.ptr(wp), L.ptr(bp), L.ptr(g1), L.ptr(be1), L.ptr(ln), L.ptr(w1), L.ptr(b1), L.ptr(w2),  # noqa: E731
                       L.ptr(b2), L.ptr(xx), L.ptr(gamma) if with_ln else None, L.ptr(beta) if with_ln else None,
                       L.ptr(ln) if with_ln else None, M, None, L.stream_ptr())
    _sync_check(L.lib().pd_op_proj_ffn_fused(*args(xo)))
    assert rel_err(xo, ref_x) < 3e-3   # bf16 roundings of ln1 / mid may flip at ties between the implementations
    if with_ln:
        assert rel_err(ln, ref_ln) < 1e-2
    xo2 = x.clone()
    _sync_check(L.lib().pd_op_proj_ffn_fused(*args(xo2)))
    assert torch.equal(xo, xo2)


@pytest.mark.parametrize("B,D,H,W,C,N,P,with_ln", [
    (4, 13, 16, 16, 256, 256, 36, True),    # shipped level-0 Conv3d: 26 tiles x 108 k-blocks per sample -> 36 ranges of 78
    (2, 13, 8, 8, 512, 512, 36, False),     # shipped level-1 Conv3d: 14 tiles x 216 k-blocks -> 36 ranges of 84 (4 parts)
    (1, 13, 16, 16, 128, 256, 30, False),   # ragged cut: 26 x 54 units in 30 ranges
    (3, 6, 16, 16, 128, 256, 13, True),
])
def test_conv3d_streamk(B, D, H, W, C, N, P, with_ln):
    """Stream-K schedule of the implicit-GEMM conv vs torch conv3d; bias + per-sample row vector + in-place residual
    (+ fused LayerNorm); deterministic and identical for a sub-batch (batch invariance)."""
    x = _randn(B, D, H, W, C, seed=21).bfloat16()
    w = _randn(N, C, 3, 3, 3, seed=22, scale=(27 * C) ** -0.5)
    wp = _pack_conv(w, C)
    bias = _randn(N, seed=23)
    rv = _randn(B, N, seed=24)
    res = _randn(B, D, H, W, N, seed=25)
    gamma, beta = 1 + 0.1 * _randn(N, seed=26), 0.1 * _randn(N, seed=27)
    ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w.bfloat16().float(), bias, padding=1).permute(0, 2, 3, 4, 1)
    ref = ref + rv[:, None, None, None, :] + res

    def run(xs, rvs, ress, b):
        out = ress.clone()
        ln = torch.zeros(b, D, H, W, N, device=DEV, dtype=torch.bfloat16)
        _sync_check(L.lib().pd_op_conv_gemm_streamk(L.ptr(xs), L.ptr(wp), b, D, H, W, C, 3, 3, 3, N, L.ptr(bias), L.ptr(rvs),
                                                    L.ptr(out), L.ptr(out), L.ptr(gamma) if with_ln else None,
                                                    L.ptr(beta) if with_ln else None, L.ptr(ln) if with_ln else None, P,
                                                    L.stream_ptr()))
        return out, ln

    out, ln = run(x, rv, res, B)
    assert rel_err(out, ref) < 3e-5
    if with_ln:
        assert rel_err(ln, F.layer_norm(ref, (N,), gamma, beta, 1e-5)) < 6e-3
    out2, _ = run(x, rv, res, B)
    assert torch.equal(out, out2)
    if B > 1:   # the last sample alone: same bits (the cut depends on the layer shape only)
        o1, _ = run(x[-1:].contiguous(), rv[-1:].contiguous(), res[-1:].contiguous(), 1)
        assert torch.equal(o1[0], out[-1])


@pytest.mark.parametrize("B,D,H,W,C,N,P", [
    (2, 13, 16, 16, 256, 256, 36),    # level-0 resblock conv1 (stream-K): 8 channels per group
    (2, 13, 8, 8, 512, 512, 36),      # level-1 (ragged last tile: 832 rows = 6.5 tiles): 16 channels per group
    (3, 4, 8, 8, 128, 256, 0),        # plain kernel (gemm_tc_kernel<256, 4, GN>)
    (2, 13, 8, 8, 64, 512, 0),        # plain kernel, two n-tiles, ragged rows
])
def test_conv_fused_groupnorm_statistics(B, D, H, W, C, N, P):
    """GroupNorm (sum, sum of squares) tables accumulated by the conv epilogue == statistics of the tensor it wrote."""
    x = _randn(B, D, H, W, C, seed=31).bfloat16()
    w = _randn(N, C, 3, 3, 3, seed=32, scale=(27 * C) ** -0.5)
    wp = _pack_conv(w, C)
    bias = _randn(N, seed=33)
    res = _randn(B, D, H, W, N, seed=34)
    out = res.clone()
    sums = torch.zeros(B, 32, 2, device=DEV, dtype=torch.float64)
    _sync_check(L.lib().pd_op_conv_gemm_gnstats(L.ptr(x), L.ptr(wp), B, D, H, W, C, 3, 3, 3, N, L.ptr(bias), L.ptr(out),
                                                L.ptr(out), L.ptr(sums), 32, P, L.stream_ptr()))
    ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w.bfloat16().float(), bias, padding=1).permute(0, 2, 3, 4, 1) + res
    assert rel_err(out, ref) < 3e-5
    g = out.double().reshape(B, -1, 32, N // 32)           # statistics of exactly what was written
    want = torch.stack([g.sum(dim=(1, 3)), (g * g).sum(dim=(1, 3))], dim=-1)
    assert torch.allclose(sums, want, rtol=2e-6, atol=1e-3)
    # and they reproduce torch's group_norm through the same mean / variance formula gn_apply uses
    n = g.shape[1] * g.shape[3]
    mean, var = sums[..., 0] / n, sums[..., 1] / n - (sums[..., 0] / n) ** 2
    gn = F.group_norm(ref.permute(0, 4, 1, 2, 3), 32, eps=1e-5).permute(0, 2, 3, 4, 1)
    mine = (out.reshape(B, -1, 32, N // 32) - mean[:, None, :, None]) / torch.sqrt(var[:, None, :, None] + 1e-5)
    assert rel_err(mine.reshape_as(gn).float(), gn) < 1e-4


def test_gemm_plain_no_epilogue_and_rowvec():
    M, K, N, samples = 512, 128, 128, 4
    a = _randn(samples * M, K, seed=5).bfloat16()
    w = _randn(N, K, seed=6, scale=K ** -0.5).bfloat16()
    rv = _randn(samples, N, seed=7)
    out = torch.empty(samples * M, N, device=DEV)
    # 4 samples of a (D=1,H=4,W=128) grid -> rowvec indexed per sample
    _sync_check(L.lib().pd_op_conv_gemm(L.ptr(a), L.ptr(w), samples, 1, 4, 128, K, 1, 1, 1, N, None, L.ptr(rv), None,
                                        L.ptr(out), None, 0, 0, L.stream_ptr()))
    ref = (a.float() @ w.float().t()).view(samples, M, N) + rv[:, None, :]
    assert rel_err(out.view(samples, M, N), ref) < 2e-5


def _pack_conv(w, cipad):
    co, ci = w.shape[:2]
    taps = w[0, 0].numel()
    out = torch.empty(co, taps, cipad, device=DEV, dtype=torch.bfloat16)
    _sync_check(L.lib().pd_op_pack_conv(L.ptr(w.contiguous()), L.ptr(out), co, ci, taps, cipad, L.stream_ptr()))
    return out


@pytest.mark.parametrize("B,D,H,W,C,N,k", [
    (2, 13, 16, 16, 64, 64, (3, 3, 3)),     # level-0 geometry: tile = 8 x 16 half plane
    (2, 13, 8, 8, 128, 128, (3, 3, 3)),     # level-1 geometry: tile = 2 frames, ragged last tile (832 = 6.5 x 128)
    (3, 1, 16, 16, 64, 64, (1, 3, 3)),      # VAE mid-block conv2d
    (2, 1, 32, 32, 64, 32, (1, 3, 3)),
    (1, 1, 128, 128, 64, 64, (1, 3, 3)),    # VAE full-resolution conv2d: tile = one image row
    (2, 13, 16, 16, 128, 64, (1, 1, 1)),    # 1x1x1 skip conv
    (1, 13, 16, 16, 256, 256, (3, 3, 3)),   # shipped level-0 conv
    (2, 13, 8, 8, 512, 512, (3, 3, 3)),     # shipped level-1 conv: 216 k-blocks -> deterministic split-K = 2
])
def test_conv_gemm(B, D, H, W, C, N, k):
    kt, kh, kw = k
    x = _randn(B, D, H, W, C, seed=11).bfloat16()
    w = _randn(N, C, kt, kh, kw, seed=12, scale=(C * kt * kh * kw) ** -0.5)
    wp = _pack_conv(w, C)
    bias = _randn(N, seed=13)
    temb = _randn(B, N, seed=14)
    res = _randn(B, D, H, W, N, seed=15)
    out = torch.full((B, D, H, W, N), float("nan"), device=DEV)
    _sync_check(L.lib().pd_op_conv_gemm(L.ptr(x), L.ptr(wp), B, D, H, W, C, kt, kh, kw, N, L.ptr(bias), L.ptr(temb),
                                        L.ptr(res), L.ptr(out), None, 0, 0, L.stream_ptr()))
    xr = x.float().permute(0, 4, 1, 2, 3)
    wr = w.bfloat16().float()
    ref = F.conv3d(xr, wr, bias, padding=(kt // 2, kh // 2, kw // 2)) + temb[:, :, None, None, None]
    ref = ref.permute(0, 2, 3, 4, 1) + res
    e = rel_err(out, ref)
    assert e < 3e-5, f"conv rel err {e}"


def test_conv_padded_input_channels():
    # 65 real channels padded to 128 (first_proj): padded activations/weights are zero
    B, D, H, W, Cr, Cp, N = 1, 13, 16, 16, 65, 128, 64
    xr = _randn(B, D, H, W, Cr, seed=21).bfloat16()
    x = torch.zeros(B, D, H, W, Cp, device=DEV, dtype=torch.bfloat16)
    x[..., :Cr] = xr
    w = _randn(N, Cr, 3, 3, 3, seed=22, scale=0.02)
    wp = _pack_conv(w, Cp)
    out = torch.empty(B, D, H, W, N, device=DEV)
    _sync_check(L.lib().pd_op_conv_gemm(L.ptr(x), L.ptr(wp), B, D, H, W, Cp, 3, 3, 3, N, None, None, None, L.ptr(out),
                                        None, 0, 0, L.stream_ptr()))
    ref = F.conv3d(xr.float().permute(0, 4, 1, 2, 3), w.bfloat16().float(), padding=1).permute(0, 2, 3, 4, 1)
    assert rel_err(out, ref) < 3e-5


@pytest.mark.parametrize("Fr,H,W,C,N", [(2, 32, 32, 64, 64), (1, 128, 128, 128, 128)])
def test_conv_stride2(Fr, H, W, C, N):
    x = _randn(Fr, H, W, C, seed=31)
    w = _randn(N, C, 3, 3, seed=32, scale=(9 * C) ** -0.5)
    bias = _randn(N, seed=33)
    planes = torch.empty(Fr, 4, H // 2, W // 2, C, device=DEV, dtype=torch.bfloat16)
    _sync_check(L.lib().pd_op_parity_split_cast(L.ptr(x), L.ptr(planes), Fr, H, W, C, L.stream_ptr()))
    wp = _pack_conv(w, C)
    out = torch.empty(Fr, H // 2, W // 2, N, device=DEV)
    _sync_check(L.lib().pd_op_conv_s2_gemm(L.ptr(planes), L.ptr(wp), Fr, H // 2, W // 2, C, N, L.ptr(bias), L.ptr(out),
                                           L.stream_ptr()))
    xr = F.pad(x.bfloat16().float().permute(0, 3, 1, 2), (0, 1, 0, 1))
    ref = F.conv2d(xr, w.bfloat16().float(), bias, stride=2).permute(0, 2, 3, 1)
    assert rel_err(out, ref) < 3e-5


# ------------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("S,R,C,G,eps,silu", [(2, 3328, 256, 32, 1e-5, 1), (2, 832, 512, 32, 1e-5, 1),
                                              (1, 3328, 128, 128, 1e-5, 1), (3, 16384, 128, 32, 1e-6, 1),
                                              (2, 256, 512, 32, 1e-6, 0), (2, 3328, 64, 32, 1e-5, 1)])
def test_group_norm(S, R, C, G, eps, silu):
    x = _randn(S, R, C, seed=41) * 2 + 0.5
    gamma = 1 + 0.1 * _randn(C, seed=42)
    beta = 0.1 * _randn(C, seed=43)
    y = torch.empty(S, R, C, device=DEV, dtype=torch.bfloat16)
    _sync_check(L.lib().pd_op_group_norm(L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(y), S, R, C, G, ctypes.c_float(eps),
                                         silu, L.stream_ptr()))
    ref = F.group_norm(x.permute(0, 2, 1), G, gamma, beta, eps).permute(0, 2, 1)
    if silu:
        ref = F.silu(ref)
    assert rel_err(y, ref) < 6e-3  # bf16 output rounding


@pytest.mark.parametrize("P,C", [(3328, 256), (832, 512), (100, 1024), (64, 2048), (77, 64), (50, 128)])
def test_layer_norm(P, C):
    x = _randn(P, C, seed=51) * 3 - 1
    gamma = 1 + 0.1 * _randn(C, seed=52)
    beta = 0.1 * _randn(C, seed=53)
    y = torch.empty(P, C, device=DEV, dtype=torch.bfloat16)
    _sync_check(L.lib().pd_op_layer_norm(L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(y), P, C, ctypes.c_float(1e-5),
                                         L.stream_ptr()))
    ref = F.layer_norm(x, (C,), gamma, beta, 1e-5)
    assert rel_err(y, ref) < 6e-3


def test_patch_merge_ln():
    BT, H, W, C = 26, 16, 16, 256
    x = _randn(BT, H, W, C, seed=61)
    gamma = 1 + 0.1 * _randn(4 * C, seed=62)
    beta = 0.1 * _randn(4 * C, seed=63)
    y = torch.empty(BT, H // 2, W // 2, 4 * C, device=DEV, dtype=torch.bfloat16)
    _sync_check(L.lib().pd_op_patch_merge_ln(L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(y), BT, H, W, C,
                                             ctypes.c_float(1e-5), L.stream_ptr()))
    # reference cuboid_transformer.py:286-294
    xr = x.reshape(BT, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(BT, H // 2, W // 2, 4 * C)
    ref = F.layer_norm(xr, (4 * C,), gamma, beta, 1e-5)
    assert rel_err(y, ref) < 6e-3


# ------------------------------------------------------------------------------------------------ attention
def _axial_ref(qkv, table, B, T, H, W, C, heads, axis):
    hd = C // heads
    q, k, v = qkv.float().view(B, T, H, W, 3, heads, hd).unbind(4)  # (B,T,H,W,heads,hd)
    dim = 1 + axis
    Lx = (T, H, W)[axis]
    q, k, v = (t.movedim(dim, 4) for t in (q, k, v))  # (B, o1, o2, heads, L, hd)
    s = (q * hd ** -0.5) @ k.transpose(-1, -2)
    idx = torch.arange(Lx, device=qkv.device)
    rel = idx[:, None] - idx[None, :] + Lx - 1
    s = s + table[rel].permute(2, 0, 1)  # (heads, L, L)
    o = torch.softmax(s, -1) @ v  # (B,o1,o2,heads,L,hd)
    o = o.movedim(4, dim)  # back to (B,T,H,W,heads,hd)
    return o.reshape(B, T, H, W, C)


@pytest.mark.parametrize("T,H,W,C,heads", [(13, 16, 16, 256, 4), (13, 8, 8, 512, 4), (13, 16, 16, 64, 4),
                                           (6, 16, 16, 128, 4)])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_axial_attention(T, H, W, C, heads, axis):
    B = 2
    qkv = _randn(B, T, H, W, 3 * C, seed=71).bfloat16()
    Lx = (T, H, W)[axis]
    table = _randn(2 * Lx - 1, heads, seed=72)
    out = torch.empty(B, T, H, W, C, device=DEV, dtype=torch.bfloat16)
    _sync_check(L.lib().pd_op_axial_attention(L.ptr(qkv), L.ptr(table), L.ptr(out), B, T, H, W, C, heads, axis,
                                              L.stream_ptr()))
    ref = _axial_ref(qkv, table, B, T, H, W, C, heads, axis)
    assert rel_err(out, ref) < 8e-3


@pytest.mark.parametrize("B,T,H,W,C,heads", [(4, 13, 16, 16, 256, 4), (1, 13, 16, 16, 256, 4), (4, 13, 8, 8, 512, 4),
                                             (2, 13, 16, 16, 64, 4), (2, 6, 16, 16, 128, 4), (3, 5, 12, 10, 128, 2),
                                             (2, 13, 8, 8, 128, 4)])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_qkv_attn_fused(B, T, H, W, C, heads, axis):
    """Fused QKV projection + axial attention core (csrc/qkv_attn.cu) against (1) the pair of kernels it replaces - the
    QKV GEMM with bf16 output and axial_attention: bit-identical - and (2) a torch fp32 reference of the layer's core
    (cuboid_transformer.py:812-861, 949) on the bf16-rounded q|k|v."""
    ln = _randn(B, T, H, W, C, seed=171).bfloat16()
    wqkv = _randn(3 * C, C, seed=172, scale=C ** -0.5).bfloat16()
    Lx = (T, H, W)[axis]
    table = _randn(2 * Lx - 1, heads, seed=173)
    out = torch.full((B, T, H, W, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    _sync_check(L.lib().pd_op_qkv_attn(L.ptr(ln), L.ptr(wqkv), L.ptr(table), L.ptr(out), B, T, H, W, C, heads, axis, None,
                                       L.stream_ptr()))
    M = B * T * H * W
    qkv = torch.empty(M, 3 * C, device=DEV, dtype=torch.bfloat16)
    _sync_check(L.lib().pd_op_conv_gemm(L.ptr(ln), L.ptr(wqkv), 1, 1, 1, M, C, 1, 1, 1, 3 * C, None, None, None, None,
                                        L.ptr(qkv), 0, 0, L.stream_ptr()))
    pair = torch.empty_like(out)
    _sync_check(L.lib().pd_op_axial_attention(L.ptr(qkv), L.ptr(table), L.ptr(pair), B, T, H, W, C, heads, axis,
                                              L.stream_ptr()))
    assert torch.isfinite(out.float()).all()
    assert torch.equal(out, pair), f"max diff vs the unfused pair {(out.float() - pair.float()).abs().max().item()}"
    ref = _axial_ref(qkv.view(B, T, H, W, 3 * C), table, B, T, H, W, C, heads, axis)
    assert rel_err(out, ref) < 8e-3


# ------------------------------------------------------------------------------------------------ small ops
def test_sampler_update():
    n = 4 * 6 * 16 * 16 * 64
    z, eps, noise, guide = (_randn(n, seed=s) for s in (81, 82, 83, 84))
    coef = torch.tensor([1.3, 0.7, 0.4, 0.55, 0.1, 0.2, 0.05, 0.0], device=DEV)
    z0 = z.clone()
    _sync_check(L.lib().pd_op_sampler_update(L.ptr(z), L.ptr(eps), L.ptr(noise), L.ptr(guide), L.ptr(coef),
                                             ctypes.c_int64(n), L.stream_ptr()))
    x0 = 1.3 * z0 - 0.7 * eps
    ref = 0.4 * x0 + 0.55 * z0 + 0.1 * eps - 0.05 * guide + 0.2 * noise
    assert (z - ref).abs().max().item() < 1e-5


def test_timestep_embedding_and_small_linear():
    B, dim = 4, 256
    t = torch.tensor([0, 1, 500, 999], device=DEV, dtype=torch.int64)
    out = torch.empty(B, dim, device=DEV)
    _sync_check(L.lib().pd_op_timestep_embedding(L.ptr(t), L.ptr(out), B, dim, L.stream_ptr()))
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half).to(DEV)
    args = t[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1)
    assert (out - ref).abs().max().item() < 2e-4
    w = _randn(1024, dim, seed=91, scale=dim ** -0.5)
    b = _randn(1024, seed=92)
    y = torch.empty(B, 1024, device=DEV)
    _sync_check(L.lib().pd_op_small_linear(L.ptr(ref), L.ptr(w), L.ptr(b), L.ptr(y), B, dim, 1024, 1, 1, L.stream_ptr()))
    yr = F.silu(F.silu(ref) @ w.t() + b)
    assert rel_err(y, yr) < 1e-5


def test_upsample_and_pack_linear():
    x = _randn(3, 8, 8, 64, seed=95)
    y = torch.empty(3, 16, 16, 64, device=DEV, dtype=torch.bfloat16)
    _sync_check(L.lib().pd_op_upsample2x_cast(L.ptr(x), L.ptr(y), 3, 8, 8, 64, L.stream_ptr()))
    ref = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1).bfloat16()
    assert torch.equal(y, ref)
    w = _randn(48, 65, seed=96)
    p = torch.empty(48, 128, device=DEV, dtype=torch.bfloat16)
    _sync_check(L.lib().pd_op_pack_linear(L.ptr(w), L.ptr(p), 48, 65, 128, L.stream_ptr()))
    assert torch.equal(p[:, :65], w.bfloat16()) and p[:, 65:].float().abs().max().item() == 0.0
