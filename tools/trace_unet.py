"""Per-launch device timeline of one UNet forward (shipped config): a %globaltimer stamp kernel after every plan step
(pd_unet_trace_forward), aggregated by kernel class / call site. Unlike ncu's serialised cold-cache replay this runs
the real back-to-back launch sequence with warm caches; each delta includes one stamp-kernel slot (reported, ~1.3 us).

  python tools/trace_unet.py --batch 4 [--out profiles/trace_unet_b4.txt]
"""
import argparse
import collections
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import _lib as L  # noqa: E402
from prediff_b200 import weights as Wt  # noqa: E402
from prediff_b200.unet import CuboidTransformerUNet  # noqa: E402

if os.environ.get("PD_LIB_PATH"):   # A/B against an older build of the library: entry points added since are no-ops
    for _name in ("pd_unet_set_precision",):
        try:
            getattr(L.lib(), _name)
        except AttributeError:
            setattr(L.lib(), _name, lambda *a: 0)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--out", default=None)
ap.add_argument("--graph", action="store_true",
                help="capture the stamped forward into a CUDA graph and replay it: launches are then issued by the GPU "
                     "front end, not by the host (an eager trace is host-launch-bound for kernels shorter than ~5 us)")
ap.add_argument("--patterns", default="axial,axial",
                help="block_attn_patterns of the two levels (any registered name, e.g. video_swin_2x8,spatial_lg_4)")
ap.add_argument("--padding", default="zeros", choices=["zeros", "ignore"])
ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"])
ap.add_argument("--global-vectors", type=int, default=0, help="num_global_vectors (with use_global_vector_ffn; 0 = none)")
ap.add_argument("--global-self-attn", action="store_true")
args = ap.parse_args()
cfg = Wt.UNetConfig(patterns=tuple(args.patterns.split(",")), padding_type=args.padding, num_global_vectors=args.global_vectors,
                    use_global_self_attn=args.global_self_attn)
B = args.batch
unet = CuboidTransformerUNet([cfg.t_in, cfg.h, cfg.w, cfg.c], [cfg.t_out, cfg.h, cfg.w, cfg.c], base_units=cfg.base_units,
                             depth=list(cfg.depth), num_heads=cfg.num_heads, block_attn_patterns=list(cfg.patterns),
                             padding_type=cfg.padding_type, max_batch=B, precision=args.precision,
                             num_global_vectors=cfg.num_global_vectors, use_global_vector_ffn=cfg.use_global_vector_ffn,
                             use_global_self_attn=cfg.use_global_self_attn)
unet.load_state_dict({k: torch.from_numpy(v) for k, v in Wt.seeded_state_dict(Wt.unet_param_spec(cfg), 1001).items()},
                     strict=False)
rng = np.random.Generator(np.random.PCG64(1))
x = torch.from_numpy(rng.standard_normal((B, cfg.t_out, cfg.h, cfg.w, cfg.c), dtype=np.float32)).cuda()
cond = torch.from_numpy(rng.standard_normal((B, cfg.t_in, cfg.h, cfg.w, cfg.c), dtype=np.float32)).cuda()
t = torch.full((B,), 500, device="cuda", dtype=torch.int64)
out = torch.empty_like(x)
SLOTS = 2048
ns = torch.zeros(SLOTS, device="cuda", dtype=torch.int64)
labels = ctypes.create_string_buffer(1 << 16)
n = 0


def traced():
    global n
    n = L.lib().pd_unet_trace_forward(unet.handle, L.ptr(x), L.ptr(t), L.ptr(cond), L.ptr(out), B, L.stream_ptr(), L.ptr(ns),
                                      SLOTS, labels, len(labels))
    if n < 0:
        L.check(n)


for _ in range(4):
    traced()
    torch.cuda.synchronize()
if args.graph:
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        traced()
    for _ in range(4):
        g.replay()
        torch.cuda.synchronize()
lab = labels.value.decode().split("\n")[:n]
s = ns.cpu().numpy()[:n + 1]
d = np.diff(s).astype(np.float64) * 1e-3   # us per step (incl. one stamp slot)
# stamp slot: an empty placeholder step would show it; estimate with the smallest delta
slot = float(np.min(d))
agg = collections.OrderedDict()
for name, us in zip(lab, d):
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us - slot
total = sum(v[1] for v in agg.values())
lines = [f"UNet forward, batch {B}, {args.precision} operands, patterns {args.patterns} / {args.padding}{' (CUDA-graph replay)' if args.graph else ''}: {n} plan steps, {(s[-1] - s[0]) * 1e-3:.1f} us wall with stamps, "
         f"{total:.1f} us after removing {n} stamp slots of {slot:.2f} us",
         f"{'site':28s} {'n':>4s} {'avg us':>8s} {'total us':>9s} {'share':>6s}"]
for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{name:28s} {cnt:4d} {us / cnt:8.2f} {us:9.1f} {100 * us / total:5.1f}%")
txt = "\n".join(lines)
print(txt)
if args.out:
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    open(args.out, "w").write(txt + "\n")
