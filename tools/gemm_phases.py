"""Per-phase cycle stamps of one CTA of the tcgen05 GEMM for the UNet's shapes (B=4) + CUDA-event kernel times."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import _lib as L  # noqa: E402

L.init()
dev = "cuda"
BLOCKS = (0, 25, 1000, 1025) if os.environ.get("PD_PHASE_SPLIT") else (0, 25, 50)   # x + 1000 * z
COLD = bool(os.environ.get("PD_PHASE_COLD"))   # weights cold in L2 (as in the real step: 274 MB of weights > L2)
FLUSH = torch.zeros(96 * 1024 * 1024, device=dev) if COLD else None
names = ["setup", "first_tile", "mainloop", "acc_ready", "first_chunk", "epilogue", "drain", "exit"]


def run(tag, samples, D, H, W, C, k, N, res, bf16_out, act, bn=0):
    kt, kh, kw = k
    M = samples * D * H * W
    a = torch.randn(M, C, device=dev).bfloat16()
    w = (torch.randn(N, kt * kh * kw * C, device=dev) * 0.02).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16 if bf16_out else torch.float32)
    stamps = torch.zeros(16, device=dev, dtype=torch.int64)
    args = lambda blk: (L.ptr(a), L.ptr(w), samples, D, H, W, C, kt, kh, kw, N, L.ptr(bias), L.ptr(out) if res else None,
                        None if bf16_out else L.ptr(out), L.ptr(out) if bf16_out else None, act, bn, blk, L.ptr(stamps),
                        L.stream_ptr())
    for _ in range(3):
        L.check(L.lib().pd_op_conv_gemm_phases(*args(0)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        L.check(L.lib().pd_op_conv_gemm_phases(*args(0)))
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 20
    flops = 2.0 * M * N * C * kt * kh * kw
    line = f"{tag:28s} {us:7.1f} us/launch (back-to-back) {flops / us * 1e-6:7.1f} TF/s |"
    t_first = None
    for blk in BLOCKS:
        stamps.zero_()
        if COLD:   # evict the weights (and everything else) from the 126 MB L2, then re-touch the activations only
            FLUSH.add_(1.0)
            a.add_(0)
            out.add_(0)
        L.check(L.lib().pd_op_conv_gemm_phases(*args(blk)))
        torch.cuda.synchronize()
        s = stamps.cpu().tolist()
        ns = s[10] - s[9]
        if bf16_out and s[2] and not s[7]:   # persistent kernel: entry, setup, (epilogue begin, end) x tiles, exit
            line += f" cta{blk}(persistent): setup={s[1]-s[0]} " + " ".join(
                f"t{i}:wait_acc@{s[2+2*i]-s[0]} epi={s[3+2*i]-s[2+2*i]}" for i in range(3) if s[2 + 2 * i]) + \
                f" total={s[8]-s[0]} cyc = {ns} ns |"
            stamps.zero_()
            continue
        d = [s[i + 1] - s[i] for i in range(8)]
        line += f" cta{blk}: " + " ".join(f"{n}={v}" for n, v in zip(names, d)) + \
            f" total={s[8] - s[0]} cyc = {ns} ns -> {(s[8] - s[0]) / max(ns, 1):.3f} GHz |"
    print(line)


# level 0 (B=4): P = 13312 tokens, C=256
run("L0 qkv   256->768 bf16", 1, 1, 1, 13312, 256, (1, 1, 1), 768, False, True, 0)
run("L0 proj  256->256 f32+res", 1, 1, 1, 13312, 256, (1, 1, 1), 256, True, False, 0)
run("L0 ffn1  256->1024 gelu", 1, 1, 1, 13312, 256, (1, 1, 1), 1024, False, True, 1)
run("L0 ffn2  1024->256 f32+res", 1, 1, 1, 13312, 1024, (1, 1, 1), 256, True, False, 0)
run("L0 conv3d 256->256", 4, 13, 16, 16, 256, (3, 3, 3), 256, True, False, 0)
# level 1: P = 3328 tokens, C=512
run("L1 qkv   512->1536 bf16", 1, 1, 1, 3328, 512, (1, 1, 1), 1536, False, True, 0)
run("L1 proj  512->512 f32+res", 1, 1, 1, 3328, 512, (1, 1, 1), 512, True, False, 0)
run("L1 ffn1  512->2048 gelu", 1, 1, 1, 3328, 512, (1, 1, 1), 2048, False, True, 1)
run("L1 ffn2  2048->512 f32+res", 1, 1, 1, 3328, 2048, (1, 1, 1), 512, True, False, 0)
run("L1 conv3d 512->512", 4, 13, 8, 8, 512, (3, 3, 3), 512, True, False, 0)
run("L1 conv3d 512->512 bn256", 4, 13, 8, 8, 512, (3, 3, 3), 512, True, False, 0, 256)
if os.environ.get("PD_PHASE_SPLIT"):
    names_blocks = (0, 25, 1000, 1025)

