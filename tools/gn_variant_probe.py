"""Same stream-K / plain conv launch with and without the fused GroupNorm-statistics epilogue (run under ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import _lib as L
L.init()
B, D, H, W, C, N = 4, 13, 16, 16, 256, 256
x = torch.randn(B, D, H, W, C, device="cuda").bfloat16()
w = (torch.randn(N, 27 * C, device="cuda") * 0.02).bfloat16()
bias = torch.randn(N, device="cuda")
out = torch.zeros(B, D, H, W, N, device="cuda")
part = torch.zeros(B, 32, 2, device="cuda", dtype=torch.float64)
for P in (36, 0):
    for gp in (None, part):
        for _ in range(3):
            L.check(L.lib().pd_op_conv_gemm_gnstats(L.ptr(x), L.ptr(w), B, D, H, W, C, 3, 3, 3, N, L.ptr(bias), L.ptr(out),
                                                    L.ptr(out), L.ptr(gp) if gp is not None else None, 32, P, L.stream_ptr()))
        torch.cuda.synchronize()
