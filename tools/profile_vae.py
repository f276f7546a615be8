"""Device time of AutoencoderKL.encode (28 frames) / decode (24 frames) at the benchmark's batch (4 forecasts): CUDA events
around repeated calls + the per-class split of pd_vae_profile (if exported)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prediff_b200 import weights as Wt  # noqa: E402
from prediff_b200.vae import AutoencoderKL  # noqa: E402

vcfg = Wt.VAEConfig()
vae = AutoencoderKL(block_out_channels=vcfg.block_out_channels, layers_per_block=vcfg.layers_per_block,
                    latent_channels=vcfg.latent_channels, sample_size=(vcfg.h, vcfg.w), max_frames=28)
vae.load_state_dict({k: torch.from_numpy(v) for k, v in Wt.seeded_state_dict(Wt.vae_param_spec(vcfg), 2002).items()},
                    strict=True)
rng = np.random.Generator(np.random.PCG64(3))
x = torch.from_numpy(rng.random((28, 1, vcfg.h, vcfg.w), dtype=np.float32)).cuda()
z = torch.from_numpy(rng.standard_normal((24, vcfg.latent_channels, vcfg.h // 8, vcfg.w // 8), dtype=np.float32)).cuda()


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if "--profile" in sys.argv:   # one encode + one decode between cudaProfilerStart/Stop: the target of the ncu passes
    for _ in range(2):
        vae.encode(x)
        vae.decode(z)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    vae.encode(x)
    vae.decode(z)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("profiled one encode (28 frames) + one decode (24 frames)")
    sys.exit(0)
enc = timeit(lambda: vae.encode(x))
dec = timeit(lambda: vae.decode(z))
print(f"encode 28 frames: {enc:.3f} ms ({28 * 67.99 / enc:.0f} TFLOP/s)   decode 24 frames: {dec:.3f} ms "
      f"({24 * 155.21 / dec:.0f} TFLOP/s)")
