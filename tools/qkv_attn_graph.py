"""Device time per launch of the fused QKV + axial attention kernel inside a CUDA graph of N back-to-back launches (no host
launch cost), against the pair of kernels it replaces."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import _lib as L  # noqa: E402

L.init()
dev = "cuda"
B = int(os.environ.get("B", 4))
N = 20


def graph_time(fn):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(N):
                fn()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / (10 * N)


for (T, H, W, C, heads) in [(13, 16, 16, 256, 4), (13, 8, 8, 512, 4)]:
    ln = torch.randn(B, T, H, W, C, device=dev).bfloat16()
    wqkv = (torch.randn(3 * C, C, device=dev) * C ** -0.5).bfloat16()
    out = torch.empty(B, T, H, W, C, device=dev, dtype=torch.bfloat16)
    M = B * T * H * W
    qkv = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
    for axis in (0, 1, 2):
        Lx = (T, H, W)[axis]
        table = torch.randn(2 * Lx - 1, heads, device=dev)

        def fused():
            L.check(L.lib().pd_op_qkv_attn(L.ptr(ln), L.ptr(wqkv), L.ptr(table), L.ptr(out), B, T, H, W, C, heads, axis, None,
                                           L.stream_ptr()))

        def pair():
            L.check(L.lib().pd_op_conv_gemm(L.ptr(ln), L.ptr(wqkv), 1, 1, 1, M, C, 1, 1, 1, 3 * C, None, None, None, None,
                                            L.ptr(qkv), 0, 0, L.stream_ptr()))
            L.check(L.lib().pd_op_axial_attention(L.ptr(qkv), L.ptr(table), L.ptr(out), B, T, H, W, C, heads, axis,
                                                  L.stream_ptr()))

        print(f"C={C} axis={axis} (L={Lx}), batch {B}: fused {graph_time(fused):.2f} us | QKV GEMM + axial_attention "
              f"{graph_time(pair):.2f} us per launch (graph of {N})", flush=True)
