"""One launch of the fused QKV + axial attention kernel at the batch-4 level-0 shape (ncu target)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import _lib as L  # noqa: E402

L.init()
B, T, H, W, C, heads = 4, 13, 16, 16, 256, 4
axis = int(os.environ.get("AXIS", 2))
ln = torch.randn(B, T, H, W, C, device="cuda").bfloat16()
wqkv = (torch.randn(3 * C, C, device="cuda") * C ** -0.5).bfloat16()
out = torch.empty(B, T, H, W, C, device="cuda", dtype=torch.bfloat16)
table = torch.randn(2 * (T, H, W)[axis] - 1, heads, device="cuda")
for _ in range(3):
    L.check(L.lib().pd_op_qkv_attn(L.ptr(ln), L.ptr(wqkv), L.ptr(table), L.ptr(out), B, T, H, W, C, heads, axis, None,
                                   L.stream_ptr()))
torch.cuda.synchronize()
