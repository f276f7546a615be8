"""One eager UNet forward (shipped config) bracketed by cudaProfilerStart/Stop - the target of the ncu passes:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_unet.py --batch 4
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -c 6 \
      -o gpurun_out/prof_gemm python tools/profile_unet.py --batch 4
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import weights as Wt  # noqa: E402
from prediff_b200.unet import CuboidTransformerUNet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--tiny", action="store_true")
ap.add_argument("--patterns", default="axial,axial", help="block_attn_patterns of the two levels")
ap.add_argument("--padding", default="zeros", choices=["zeros", "ignore"])
ap.add_argument("--depth", default=None, help="e.g. 1,1 (default: the config's)")
args = ap.parse_args()
import dataclasses  # noqa: E402
cfg = dataclasses.replace(Wt.TINY_UNET if args.tiny else Wt.UNetConfig(), patterns=tuple(args.patterns.split(",")),
                          padding_type=args.padding)
if args.depth:
    cfg = dataclasses.replace(cfg, depth=tuple(int(v) for v in args.depth.split(",")))
B = args.batch
unet = CuboidTransformerUNet([cfg.t_in, cfg.h, cfg.w, cfg.c], [cfg.t_out, cfg.h, cfg.w, cfg.c], base_units=cfg.base_units,
                             depth=list(cfg.depth), num_heads=cfg.num_heads, block_attn_patterns=list(cfg.patterns),
                             padding_type=cfg.padding_type, max_batch=B)
unet.load_state_dict({k: torch.from_numpy(v) for k, v in Wt.seeded_state_dict(Wt.unet_param_spec(cfg), 1001).items()},
                     strict=False)
rng = np.random.Generator(np.random.PCG64(1))
x = torch.from_numpy(rng.standard_normal((B, cfg.t_out, cfg.h, cfg.w, cfg.c), dtype=np.float32)).cuda()
cond = torch.from_numpy(rng.standard_normal((B, cfg.t_in, cfg.h, cfg.w, cfg.c), dtype=np.float32)).cuda()
t = torch.full((B,), 500, device="cuda", dtype=torch.int64)
for _ in range(2):
    unet(x, t, cond)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
unet(x, t, cond)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one forward, batch", B)
