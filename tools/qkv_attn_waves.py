"""Entry / exit time (%globaltimer) and SM of EVERY CTA of one launch of the fused QKV + axial attention kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import _lib as L  # noqa: E402

L.init()
dev = "cuda"
B = int(os.environ.get("B", 4))
T, H, W, C, heads = [int(v) for v in os.environ.get("SHAPE", "13,8,8,512,4").split(",")]
os.environ["PD_QKV_DBG_CTA"] = "-1,-1"
for axis in (0, 1, 2):
    ln = torch.randn(B, T, H, W, C, device=dev).bfloat16()
    wqkv = (torch.randn(3 * C, C, device=dev) * C ** -0.5).bfloat16()
    out = torch.empty(B, T, H, W, C, device=dev, dtype=torch.bfloat16)
    Lx = (T, H, W)[axis]
    table = torch.randn(2 * Lx - 1, heads, device=dev)
    st = torch.zeros(32 + 3 * 1024, device=dev, dtype=torch.int64)
    for _ in range(3):
        st.zero_()
        torch.cuda.synchronize()
        L.check(L.lib().pd_op_qkv_attn(L.ptr(ln), L.ptr(wqkv), L.ptr(table), L.ptr(out), B, T, H, W, C, heads, axis, L.ptr(st),
                                       L.stream_ptr()))
        torch.cuda.synchronize()
    s = st.cpu()[32:].view(-1, 3)
    s = s[s[:, 0] > 0]
    t0 = s[:, 0].min().item()
    ent, ex, sm = (s[:, 0] - t0).tolist(), (s[:, 1] - t0).tolist(), s[:, 2].tolist()
    print(f"axis {axis}: {len(ent)} CTAs on {len(set(sm))} distinct SMs; kernel span {max(ex)} ns; entries: "
          f"{sorted(ent)[:3]} ... {sorted(ent)[-3:]}; late entries (> 2 us): {sum(e > 2000 for e in ent)}")
    from collections import Counter
    multi = {k: v for k, v in Counter(sm).items() if v > 1}
    print("   SMs that ran more than one CTA:", multi)
