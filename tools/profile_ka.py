"""One knowledge-alignment guidance evaluation (KA forward + input-gradient backward, shipped config) bracketed by
cudaProfilerStart/Stop - the target of the ncu passes:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches_ka.csv python tools/profile_ka.py --batch 4
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import weights as Wt  # noqa: E402
from prediff_b200.alignment import SEVIRAvgIntensityAlignment  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
args = ap.parse_args()
cfg = Wt.KAConfig()
B = args.batch
al = SEVIRAvgIntensityAlignment(guide_scale=cfg.guide_scale, model_args=dict(
    input_shape=[cfg.t, cfg.h, cfg.w, cfg.c], base_units=cfg.base_units, depth=list(cfg.depth), block_attn_patterns="axial",
    num_heads=cfg.num_heads, pool="attention", readout_seq=True, out_len=cfg.t, max_batch=B))
al.model.load_state_dict({k: torch.from_numpy(v) for k, v in Wt.seeded_state_dict(Wt.ka_param_spec(cfg), 3003).items()},
                         strict=False)
rng = np.random.Generator(np.random.PCG64(1))
zt = torch.from_numpy(rng.standard_normal((B, cfg.t, cfg.h, cfg.w, cfg.c), dtype=np.float32)).cuda()
t = torch.full((B,), 500, device="cuda", dtype=torch.int64)
tgt = torch.full((B, 1), 0.3, device="cuda")
for _ in range(2):
    al.get_mean_shift(zt, t, avg_x_gt=tgt)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
g = al.get_mean_shift(zt, t, avg_x_gt=tgt)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one KA guidance evaluation, batch", B, "grad absmax", g.abs().max().item())
