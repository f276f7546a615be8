"""Per-CTA timeline of the fused QKV + axial attention kernel: entry / exit (%globaltimer, relative to the earliest entry seen),
SM id and the phase stamps of a sample of CTAs. PD_QKV_DBG_CTA selects the stamping CTA per launch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import _lib as L  # noqa: E402

L.init()
dev = "cuda"
B = int(os.environ.get("B", 4))
T, H, W, C, heads = [int(v) for v in os.environ.get("SHAPE", "13,8,8,512,4").split(",")]
axis = int(os.environ.get("AXIS", 1))
ln = torch.randn(B, T, H, W, C, device=dev).bfloat16()
wqkv = (torch.randn(3 * C, C, device=dev) * C ** -0.5).bfloat16()
out = torch.empty(B, T, H, W, C, device=dev, dtype=torch.bfloat16)
Lx = (T, H, W)[axis]
table = torch.randn(2 * Lx - 1, heads, device=dev)
st = torch.zeros(32, device=dev, dtype=torch.int64)
ctas = [c.split(",") for c in os.environ.get("CTAS", "0,0;1,0;13,0;27,0;0,1;13,2;27,3").split(";")]
for cx, cy in ctas:
    os.environ["PD_QKV_DBG_CTA"] = f"{cx},{cy}"
    steady = int(os.environ.get("STEADY", 0))   # > 0: the stamped launch sits in the middle of 2 * STEADY unsynchronised ones
    for _ in range(3):
        st.zero_()
        torch.cuda.synchronize()
        for _ in range(steady):
            L.check(L.lib().pd_op_qkv_attn(L.ptr(ln), L.ptr(wqkv), L.ptr(table), L.ptr(out), B, T, H, W, C, heads, axis, None,
                                           L.stream_ptr()))
        L.check(L.lib().pd_op_qkv_attn(L.ptr(ln), L.ptr(wqkv), L.ptr(table), L.ptr(out), B, T, H, W, C, heads, axis, L.ptr(st),
                                       L.stream_ptr()))
        for _ in range(steady):
            L.check(L.lib().pd_op_qkv_attn(L.ptr(ln), L.ptr(wqkv), L.ptr(table), L.ptr(out), B, T, H, W, C, heads, axis, None,
                                           L.stream_ptr()))
        torch.cuda.synchronize()
    s = st.cpu().tolist()
    rel = {i: s[i] - s[0] for i in (1, 16, 17, 2, 3, 4, 5, 10)}
    print(f"CTA ({cx},{cy}) sm {s[11]:3d}: life {s[9] - s[8]} ns | dep wait +{rel[1]} operands +{rel[16]} issued +{rel[17]} "
          f"acc +{rel[2]} staged +{rel[3]} lines +{rel[4]} heads done +{rel[5]} exit +{rel[10]}")
