"""Writes the plan-step labels of one UNet forward (launch order) - pairs an ncu launch list with call sites."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import _lib as L  # noqa: E402
from prediff_b200 import weights as Wt  # noqa: E402
from prediff_b200.unet import CuboidTransformerUNet  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = Wt.UNetConfig()
unet = CuboidTransformerUNet([cfg.t_in, cfg.h, cfg.w, cfg.c], [cfg.t_out, cfg.h, cfg.w, cfg.c], base_units=cfg.base_units,
                             depth=list(cfg.depth), num_heads=cfg.num_heads, max_batch=B)
unet.load_state_dict({k: torch.from_numpy(v) for k, v in Wt.seeded_state_dict(Wt.unet_param_spec(cfg), 1001).items()},
                     strict=False)
x = torch.zeros(B, cfg.t_out, cfg.h, cfg.w, cfg.c, device="cuda")
cond = torch.zeros(B, cfg.t_in, cfg.h, cfg.w, cfg.c, device="cuda")
t = torch.full((B,), 500, device="cuda", dtype=torch.int64)
ns = torch.zeros(2048, device="cuda", dtype=torch.int64)
labels = ctypes.create_string_buffer(1 << 16)
n = L.lib().pd_unet_trace_forward(unet.handle, L.ptr(x), L.ptr(t), L.ptr(cond), L.ptr(torch.empty_like(x)), B, L.stream_ptr(),
                                  L.ptr(ns), 2048, labels, len(labels))
fl = (ctypes.c_double * 2048)()
nf = L.lib().pd_unet_step_flops(unet.handle, B, fl, 2048)
torch.cuda.synchronize()
for i, lab in enumerate(labels.value.decode().split("\n")[:n]):
    print(lab, fl[i] if i < nf else 0.0)
