"""Latency of a single CuboidTransformerUNet denoise step (BASELINE.json configs[1]: batch 1, 13 x 16 x 16 latent):
20 warm-up + 200 timed forwards bracketed by CUDA events, eager and CUDA-graph replay.

  python tools/unet_latency.py [--batch 1] [--out gpurun_out/unet_latency_b1.json]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import weights as Wt  # noqa: E402
from prediff_b200.unet import CuboidTransformerUNet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--iters", type=int, default=200)
ap.add_argument("--out", default=None)
ap.add_argument("--global-vectors", type=int, default=0, help="num_global_vectors (with the global FFN; 0 = the shipped config)")
args = ap.parse_args()
cfg, B = Wt.UNetConfig(num_global_vectors=args.global_vectors), args.batch
unet = CuboidTransformerUNet([cfg.t_in, cfg.h, cfg.w, cfg.c], [cfg.t_out, cfg.h, cfg.w, cfg.c], base_units=cfg.base_units,
                             depth=list(cfg.depth), num_heads=cfg.num_heads, max_batch=B,
                             num_global_vectors=cfg.num_global_vectors)
unet.load_state_dict({k: torch.from_numpy(v) for k, v in Wt.seeded_state_dict(Wt.unet_param_spec(cfg), 1001).items()},
                     strict=False)
rng = np.random.Generator(np.random.PCG64(1234))
x = torch.from_numpy(rng.standard_normal((B, cfg.t_out, cfg.h, cfg.w, cfg.c), dtype=np.float32)).cuda()
cond = torch.from_numpy(rng.standard_normal((B, cfg.t_in, cfg.h, cfg.w, cfg.c), dtype=np.float32)).cuda()
t = torch.full((B,), 500, device="cuda", dtype=torch.int64)


def timed(fn):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.iters


res = {"workload": f"single UNet denoise step, batch {B}, 13x16x16x64 latent (BASELINE.json configs[1])",
       "iters": args.iters, "warmup": 20, "eager_ms": timed(lambda: unet(x, t, cond))}
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    unet(x, t, cond)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = unet(x, t, cond)
    res["graph_ms"] = timed(g.replay)
res["sample_steps_per_s_graph"] = B / (res["graph_ms"] * 1e-3)
res["tflops_graph"] = B * 653.43e-3 / (res["graph_ms"] * 1e-3)
print(json.dumps(res))
if args.out:
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    open(args.out, "w").write(json.dumps(res) + "\n")
