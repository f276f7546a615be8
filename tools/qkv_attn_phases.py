"""Phase stamps of CTA (0, 0) of the fused QKV + axial attention kernel (csrc/qkv_attn.cu) at the batch-4 shapes of both
levels and all three axes, its back-to-back launch time, and the time of the pair of kernels it replaces."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import _lib as L  # noqa: E402

L.init()
dev = "cuda"
B = int(os.environ.get("B", 4))
names = {1: "dependency wait passed", 2: "head 0 accumulator complete", 3: "head 0 staged", 4: "head 0 lines done",
         5: "all heads done", 16: "MMA: first operands landed", 17: "MMA: head 0 issued"}


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / n


for (T, H, W, C, heads) in [(13, 16, 16, 256, 4), (13, 8, 8, 512, 4)]:
    ln = torch.randn(B, T, H, W, C, device=dev).bfloat16()
    wqkv = (torch.randn(3 * C, C, device=dev) * C ** -0.5).bfloat16()
    out = torch.empty(B, T, H, W, C, device=dev, dtype=torch.bfloat16)
    M = B * T * H * W
    qkv = torch.empty(M, 3 * C, device=dev, dtype=torch.bfloat16)
    for axis in (0, 1, 2):
        Lx = (T, H, W)[axis]
        table = torch.randn(2 * Lx - 1, heads, device=dev)
        st = torch.zeros(32, device=dev, dtype=torch.int64)

        def fused(stamps=None):
            L.check(L.lib().pd_op_qkv_attn(L.ptr(ln), L.ptr(wqkv), L.ptr(table), L.ptr(out), B, T, H, W, C, heads, axis,
                                           L.ptr(stamps), L.stream_ptr()))

        def pair():
            L.check(L.lib().pd_op_conv_gemm(L.ptr(ln), L.ptr(wqkv), 1, 1, 1, M, C, 1, 1, 1, 3 * C, None, None, None, None,
                                            L.ptr(qkv), 0, 0, L.stream_ptr()))
            L.check(L.lib().pd_op_axial_attention(L.ptr(qkv), L.ptr(table), L.ptr(out), B, T, H, W, C, heads, axis,
                                                  L.stream_ptr()))

        for _ in range(3):
            fused(st)
        torch.cuda.synchronize()
        s = st.cpu().tolist()
        print(f"C={C} axis={axis} (L={Lx}), batch {B}: fused {timeit(fused):.2f} us | QKV GEMM + axial_attention "
              f"{timeit(pair):.2f} us (back to back, warm, eager)")
        print("   " + " | ".join(f"{names[i]} +{s[i] - s[0]}" for i in sorted(names, key=lambda i: s[i]) if s[i]))
