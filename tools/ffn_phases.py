"""Phase stamps of CTA 0 of the fused FFN kernel (width 256, batch-4 level-0 shape)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import _lib as L  # noqa: E402

L.init()
M, C, H = int(os.environ.get("M", 13312)), 256, 1024
dev = "cuda"
ln_in = torch.randn(M, C, device=dev).bfloat16()
w1 = (torch.randn(H, C, device=dev) * C ** -0.5).bfloat16()
w2 = (torch.randn(C, H, device=dev) * H ** -0.5).bfloat16()
b1, b2 = torch.randn(H, device=dev) * 0.1, torch.randn(C, device=dev) * 0.1
x = torch.randn(M, C, device=dev)
g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
ln = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
st = torch.zeros(32, device=dev, dtype=torch.int64)
for _ in range(3):
    L.check(L.lib().pd_op_ffn_fused_phases(L.ptr(ln_in), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(x), L.ptr(g),
                                           L.ptr(b), L.ptr(ln), M, L.ptr(st), L.stream_ptr()))
torch.cuda.synchronize()
s = st.cpu().tolist()
t0 = s[0]
print("MMA warp ready @%d, A landed @%d" % (s[1] - t0, s[2] - t0))
for c in range(4):
    print(f"chunk {c}: E1 begin @{s[12 + 2 * c] - t0:6d} dur {s[13 + 2 * c] - s[12 + 2 * c]:5d} | G2 issued @{s[3 + c] - t0:6d}")
print(f"acc2 complete @{s[28] - t0}, final epilogue {s[29] - s[28]} cycles, exit @{s[30] - t0}")

# the PROJ variant (attention projection + residual + pre-norm fused in front): the level-0 stack kernel
att = torch.randn(M, C, device=dev).bfloat16()
wp = (torch.randn(C, C, device=dev) * C ** -0.5).bfloat16()
bp = torch.randn(C, device=dev) * 0.1
scratch = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
st.zero_()
for _ in range(3):
    L.check(L.lib().pd_op_proj_ffn_fused(L.ptr(att), L.ptr(wp), L.ptr(bp), L.ptr(g), L.ptr(b), L.ptr(scratch), L.ptr(w1),
                                         L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(x), L.ptr(g), L.ptr(b), L.ptr(ln), M,
                                         L.ptr(st), L.stream_ptr()))
torch.cuda.synchronize()
s = st.cpu().tolist()
t0 = s[0]
print("PROJ variant: MMA warp ready @%d, G1(0) issued @%d" % (s[1] - t0, s[2] - t0))
print("  prologue from 'ready': A landed +%d, GEMM-0 issued +%d | x + bp parked +%d, GEMM-0 complete +%d, stats +%d, A tiles written +%d"
      % tuple(s[i] - s[1] for i in (11, 20, 7, 8, 9, 10)))
for c in range(4):
    print(f"chunk {c}: E1 begin @{s[12 + 2 * c] - t0:6d} dur {s[13 + 2 * c] - s[12 + 2 * c]:5d} | G2 issued @{s[3 + c] - t0:6d}")
print(f"acc2 complete @{s[28] - t0}, final epilogue {s[29] - s[28]} cycles, exit @{s[30] - t0}")
