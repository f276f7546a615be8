"""ncu target: the global-vector kernels at the shipped UNet shapes (batch 4, K = 8 global vectors).
  ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active \
      --clock-control none --csv --log-file gpurun_out/ncu_gv.csv python tools/profile_gv.py
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import _lib as L  # noqa: E402

I3 = ctypes.c_int32 * 3
B, K, heads = 4, 8, 4
for dims, C, size in (((13, 16, 16), 256, (13, 1, 1)), ((13, 8, 8), 512, (1, 8, 1))):
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(B, *dims, 3 * C, generator=g).bfloat16().cuda()
    gq = torch.randn(B, K, 3 * C, generator=g).cuda()
    gq16 = gq.bfloat16()
    n_rel = (2 * size[0] - 1) * (2 * size[1] - 1) * (2 * size[2] - 1)
    table = torch.randn(n_rel, heads, generator=g).cuda()
    out = torch.empty(B, *dims, C, device="cuda", dtype=torch.bfloat16)
    gout = torch.empty(B, K, C, device="cuda")
    for line in (1, 1, 0):   # the line kernel (what the UNet runs for axial layers with K <= 16), then the general kernel
        L.check(L.lib().pd_op_cuboid_attention_gv2(L.ptr(qkv), L.ptr(table), None, L.ptr(gq), L.ptr(gq16), 3 * C, L.ptr(out),
                                                   L.ptr(gout), B, *dims, C, heads, I3(*size), I3(0, 0, 0), I3(0, 0, 0), 0, K, 1, line,
                                                   L.stream_ptr()))
    # the linears of one layer: global_qkv (LayerNorm fused), global_proj, global FFN
    M = B * K
    x = torch.randn(M, C, generator=g).cuda()
    gam, bet = torch.ones(C).cuda(), torch.zeros(C).cuda()
    for N, Kin, ln, act in ((3 * C, C, True, 0), (C, C, False, 0), (4 * C, C, True, 1), (C, 4 * C, False, 0)):
        W = torch.randn(N, Kin, generator=g).cuda()
        inp = torch.randn(M, Kin, generator=g).cuda()
        o = torch.empty(M, N, device="cuda")
        for _ in range(2):
            L.check(L.lib().pd_op_gv_linear(L.ptr(inp), L.ptr(gam) if ln else None, L.ptr(bet) if ln else None, L.ptr(W), None, None,
                                            L.ptr(o), None, M, Kin, N, act, L.stream_ptr()))
torch.cuda.synchronize()
print("done")
