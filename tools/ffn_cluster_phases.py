"""Phase stamps of CTA 0 of the width-512 cluster FFN kernel (batch-4 level-1 shape) + its launch time."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import _lib as L  # noqa: E402

L.init()
M, C, H = int(os.environ.get("M", 3328)), 512, 2048
dev = "cuda"
ln_in = torch.randn(M, C, device=dev).bfloat16()
w1 = (torch.randn(H, C, device=dev) * C ** -0.5).bfloat16()
w2 = (torch.randn(C, H, device=dev) * H ** -0.5).bfloat16()
b1, b2 = torch.randn(H, device=dev) * 0.1, torch.randn(C, device=dev) * 0.1
x = torch.randn(M, C, device=dev)
g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
ln = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
st = torch.zeros(32, device=dev, dtype=torch.int64)
PROJ = int(os.environ.get("PROJ", 0))   # 1: the variant with the attention projection fused in front
att = torch.randn(M, C, device=dev).bfloat16()
wp = (torch.randn(C, C, device=dev) * C ** -0.5).bfloat16()
bp = torch.randn(C, device=dev) * 0.1


def call(stamps):
    if PROJ:
        L.check(L.lib().pd_op_proj_ffn_cluster(L.ptr(att), L.ptr(wp), L.ptr(bp), L.ptr(g), L.ptr(b), L.ptr(ln_in), L.ptr(w1),
                                               L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(x), L.ptr(g), L.ptr(b), L.ptr(ln), None,
                                               32, 832, M, L.ptr(stamps), L.stream_ptr()))
    elif stamps is None:
        L.check(L.lib().pd_op_ffn_cluster(L.ptr(ln_in), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(x), L.ptr(g),
                                          L.ptr(b), L.ptr(ln), None, 32, 832, M, L.stream_ptr()))
    else:
        L.check(L.lib().pd_op_ffn_cluster_phases(L.ptr(ln_in), L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(x), L.ptr(g),
                                                 L.ptr(b), L.ptr(ln), M, L.ptr(stamps), L.stream_ptr()))


for _ in range(3):
    call(st)
torch.cuda.synchronize()
s = st.cpu().tolist()
t0 = s[0]
names = {1: "dependency wait passed", 2: "G1(0) complete", 3: "E1(0) done", 4: "G1(1) complete", 5: "E1(1) done",
         6: "partial complete", 8: "slices sent", 9: "cluster barrier", 10: "rows reduced",
         11: "LN barrier", 12: "done", 16: "MMA: first operands landed", 17: "MMA: G1(0) issued", 18: "MMA: G1(1) issued",
         19: "MMA: waits for E1(1)", 20: "MMA: E1(1) there", 21: "MMA: G2 issued"}
if PROJ:
    names.update({22: "MMA: first G0 operands landed", 23: "MMA: G0 issued", 24: "x1 complete", 25: "E0 done (LN(x1) published)", 26: "E0: x1 rows combined", 27: "E0: row sums exchanged", 28: "E0: LN(x1) store complete"})
for i in sorted(names, key=lambda i: s[i]):
    print(f"{s[i] - t0:7d}  {names[i]}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(5):
    call(None)
e0.record()
for _ in range(50):
    call(None)
e1.record()
torch.cuda.synchronize()
print("avg launch %.2f us (back to back, warm)" % (e0.elapsed_time(e1) * 1000 / 50))
