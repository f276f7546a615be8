#!/usr/bin/env python
"""Benchmark of the PreDiff sampling hot path (BASELINE.json metric: SEVIR-LR denoise-steps/sec, 50-step DDIM,
7 -> 6 frames at 128x128).

  python bench.py --gpus N --steps K --warmup W          # this repo (CUDA path), one process per GPU under torchrun
  python bench.py --impl reference --steps K --warmup W  # the reference algorithm on the host CPU (oracle port)

A bench "step" = one full 50-step DDIM loop over one batch of 4 forecasts on each GPU (config 3 of BASELINE.json:
"SEVIR-LR full 50-step DDIM p_sample_loop, batch=4 on 1xB200"), i.e. 200 UNet evaluations ("sample-steps") per
GPU per step; `value` = sample-steps/s over all GPUs with z_T / context latents resident in HBM; `e2e` = the same
metric through LatentDiffusion.sample() from pinned host frames to pinned host forecasts (VAE encode + loop +
VAE decode + the final all-gather + both copies inside the timed region). Weights are seeded random (no
checkpoints offline), data synthetic. Prints ONE JSON line on rank 0.

The headline arm runs bf16 tensor-core operands (tolerance: rel-RMS 7e-3 per step / 8e-3 after the loop vs the fp32
reference, tests/test_*_gpu.py); `tf32_arm` in the same line is the like-for-like arm for the reference's own TF32 GPU
arithmetic (kind::tf32 operands, rel-RMS <= 2e-3, tests/test_tf32_gpu.py), measured beside it and held against a
measured TF32 peak (profiles/tf32_peak_r02.json). `single_step_b1` is BASELINE.json configs[1].
"""
import ctypes
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from prediff_b200 import weights as Wt  # noqa: E402

GFLOP_PER_SAMPLE_STEP = 653.43   # algorithmic FLOPs of one UNet forward for one sample
UNET_SEED, VAE_SEED, KA_SEED = 1001, 2002, 3003
METRIC = "SEVIR-LR denoise-steps/sec (50-step DDIM, 7->6x128x128)"
UNIT = "sample-steps/s"


def np_inp(seed, *shape, uniform=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.random(shape, dtype=np.float32) if uniform else rng.standard_normal(shape, dtype=np.float32)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"tensor_tflops": d["bf16_tflops_sustained"], "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"tensor_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}   # B200_PROFILING.md fallback


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons through NVML (no process forks) at 1 Hz while a timed region runs. Only
    local rank 0 samples: in round 1 every rank forked nvidia-smi 5 times a second during the kernel-only leg, which
    alone cost 2.8 % at 8 GPUs."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index, enabled=True):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows, self.enabled, self.h = index, threading.Event(), [], enabled, None
        if enabled:
            try:
                import pynvml
                pynvml.nvmlInit()
                self.nv = pynvml
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
                self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            except Exception:
                self.h = None

    def sample(self):
        try:
            mhz = float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            try:
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                mask = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            self.rows.append((mhz, mask))
        except Exception:
            pass

    def run(self):
        if self.h is None:
            return
        while not self.stop_flag.is_set():
            self.sample()
            self.stop_flag.wait(1.0)

    def summary(self):
        self.stop_flag.set()
        if self.is_alive():
            self.join(timeout=3)
        if self.h is not None:
            self.sample()
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable" if self.enabled else "sampled on local rank 0 only"]}
        sm = sorted(r[0] for r in self.rows)
        reasons = [n for n, bit in self.REASONS if any(r[1] & bit for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.rows),
                "how": "NVML, 1 Hz, local rank 0, during the kernel-only and end-to-end timed regions"}


def run_reference_arm(args):
    """The reference's algorithm on the host CPU: the oracle port (oracle/prediff_oracle.py, pinned against the
    unmodified reference's outputs). Each step is a bounded sample of the workload: ONE denoise step of the
    batch-of-4 DDIM loop (steps are homogeneous)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import prediff_oracle as O
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    cfg = Wt.UNetConfig()
    sd = O.to_torch_sd(Wt.seeded_state_dict(Wt.unet_param_spec(cfg), UNET_SEED))
    B = args.batch
    z = torch.from_numpy(np_inp(4242, B, cfg.t_out, cfg.h, cfg.w, cfg.c))
    cond = torch.from_numpy(np_inp(4243, B, cfg.t_in, cfg.h, cfg.w, cfg.c))
    sched = O.make_schedule()
    coefs = O.ddim_coefficients(sched, args.ddim_steps, 0.0)

    def one(k):
        t, a_t, a_prev, sigma = coefs[k % len(coefs)]
        with torch.no_grad():
            eps = O.unet_forward(sd, cfg, z, torch.full((B,), t, dtype=torch.long), cond)
            z0 = (z - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
            return a_prev ** 0.5 * z0 + (1 - a_prev) ** 0.5 * eps

    for k in range(args.warmup):
        one(k)
    t0 = time.perf_counter()
    for k in range(args.steps):
        one(k)
    dt = time.perf_counter() - t0
    value = B * args.steps / dt
    sample = (f"{args.steps} of the {args.ddim_steps} denoise steps of one batch-{B} DDIM loop (UNet forward + DDIM update); the "
              "VAE encode / decode that the GPU arm's e2e includes are NOT timed here, so the e2e ratio is conservative")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic latents, seeded random weights",
        "config": {"workload": f"SEVIR-LR 50-step DDIM p_sample_loop, batch={B} (BASELINE.json configs[2])",
                   "device": "host CPU", "torch_threads": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def cpu_baseline_leg(args, budget_s=25.0):
    """Oracle (CPU port of the reference algorithm) timed on this box's host cores on a bounded sample."""
    from oracle import prediff_oracle as O
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    cfg = Wt.UNetConfig()
    sd = O.to_torch_sd(Wt.seeded_state_dict(Wt.unet_param_spec(cfg), UNET_SEED))
    B = args.batch
    z = torch.from_numpy(np_inp(4242, B, cfg.t_out, cfg.h, cfg.w, cfg.c))
    cond = torch.from_numpy(np_inp(4243, B, cfg.t_in, cfg.h, cfg.w, cfg.c))
    t = torch.full((B,), 981, dtype=torch.long)
    with torch.no_grad():
        O.unet_forward(sd, cfg, z, t, cond)  # warm-up
        n, t0 = 0, time.perf_counter()
        while n < 2 or (time.perf_counter() - t0 < budget_s and n < 8):
            O.unet_forward(sd, cfg, z, t, cond)
            n += 1
        dt = time.perf_counter() - t0
    return {"value": B * n / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} UNet denoise steps at batch {B} (of the {args.ddim_steps} x K in the GPU run), fp32, "
                      f"torch CPU with {cores} threads"}


CUBOID_GFLOP_PER_SAMPLE_STEP = 252.9   # StackCuboidSelfAttentionBlock incl. FFN


def algorithmic_bytes(site, B):
    """Bytes one launch of a call site must move if every operand is read once and every result written once (shipped
    config: 13 x 16 x 16 tokens of width 256 at level 0, 13 x 8 x 8 of width 512 at level 1; bf16 operands / weights, fp32
    residual stream). None for sites without a closed form here. Compared with the measured DRAM bytes in `kernels`:
    below it when an operand was still in the 126 MB L2, above it when something is re-read from HBM."""
    lvl = 1 if site.startswith("L1.") else 0
    M, C = B * 13 * (64 if lvl else 256), (512 if lvl else 256)
    kind = site.split(".")[-1]
    act16, act32 = M * C * 2, M * C * 4
    if kind in ("conv1", "conv2") and site[:2] in ("L0", "L1"):
        w = 27 * C * C * 2
        # conv1: a (bf16) + W -> h (fp32); conv2: a + W + residual x -> x (+ the fused bf16 LayerNorm at width 256)
        return act16 + w + act32 + (act32 + (act16 if C == 256 else 0) if kind == "conv2" else 0)
    if kind in ("proj_ffn_fused", "proj_ffn_cluster"):
        w = (C * C + 2 * C * 4 * C) * 2
        return act16 + w + 2 * act32 + act16          # att + weights + x in / out + next LayerNorm out
    if kind.startswith("qkv_attn"):
        return act16 + 3 * C * C * 2 + act16          # ln + Wqkv -> att
    if kind == "gn_apply" and site[:2] in ("L0", "L1"):
        return act32 + act16
    if kind == "ln":
        return act32 + act16
    return None


def graph_trace(unet, B, x, t, cond, out):
    """One UNet forward replayed as a CUDA graph with a %globaltimer stamp kernel after every launch.
    Returns ({site: [launches, us, flops]}, stamp_slot_us): us = raw stamp-to-stamp interval minus the stamp slot (the
    smallest interval seen: the stamp kernel's own node, ~1.8 us); flops = algorithmic FLOPs of the site's launches."""
    import collections
    from prediff_b200 import _lib as L
    slots = 2048
    ns = torch.zeros(slots, device=x.device, dtype=torch.int64)
    labels = ctypes.create_string_buffer(1 << 16)
    n = [0]

    def traced():
        n[0] = L.lib().pd_unet_trace_forward(unet.handle, L.ptr(x), L.ptr(t), L.ptr(cond), L.ptr(out), B, L.stream_ptr(),
                                             L.ptr(ns), slots, labels, len(labels))
        if n[0] < 0:
            L.check(n[0])

    traced()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        traced()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    lab = labels.value.decode().split("\n")[:n[0]]
    fl = (ctypes.c_double * slots)()
    nf = L.lib().pd_unet_step_flops(unet.handle, B, fl, slots)
    if nf < 0:
        L.check(nf)
    d = np.diff(ns.cpu().numpy()[:n[0] + 1]).astype(np.float64) * 1e-3
    slot = float(d.min())
    agg = collections.OrderedDict()
    for i, (name, us) in enumerate(zip(lab, d)):
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += us - slot
        a[2] += fl[i] if i < nf else 0.0
    return agg, slot


def load_json(rel):
    path = os.path.join(ROOT, rel)
    return json.load(open(path)) if os.path.exists(path) else None


def build_unet(ucfg, B, precision, streamk_ctas_per_sample=0):
    from prediff_b200.unet import CuboidTransformerUNet
    unet = CuboidTransformerUNet([ucfg.t_in, ucfg.h, ucfg.w, ucfg.c], [ucfg.t_out, ucfg.h, ucfg.w, ucfg.c],
                                 base_units=ucfg.base_units, depth=list(ucfg.depth), num_heads=ucfg.num_heads,
                                 block_attn_patterns="axial", max_batch=B, precision=precision,
                                 streamk_ctas_per_sample=streamk_ctas_per_sample)
    unet.load_state_dict({k: torch.from_numpy(v) for k, v in
                          Wt.seeded_state_dict(Wt.unet_param_spec(ucfg), UNET_SEED).items()}, strict=False)
    return unet


def single_step_latency(unet, ucfg, dev, iters=200, warm=20):
    """BASELINE.json configs[1]: one denoise step, batch 1, CUDA-graph replay, CUDA events."""
    x = torch.from_numpy(np_inp(1234, 1, ucfg.t_out, ucfg.h, ucfg.w, ucfg.c)).to(dev)
    cond = torch.from_numpy(np_inp(1235, 1, ucfg.t_in, ucfg.h, ucfg.w, ucfg.c)).to(dev)
    t = torch.full((1,), 500, device=dev, dtype=torch.int64)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        unet(x, t, cond)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            unet(x, t, cond)
        for _ in range(warm):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            g.replay()
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return {"workload": "single CuboidTransformerUNet denoise step, batch 1, 13x16x16 latent (BASELINE.json configs[1])",
            "iters": iters, "warmup": warm, "ms": ms, "value": 1e3 / ms, "unit": UNIT,
            "tflops": GFLOP_PER_SAMPLE_STEP / ms, "timing": "CUDA-graph replay of pd_unet_forward, CUDA events"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4, help="forecasts per GPU (BASELINE config 3: 4)")
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ka", action="store_true", help="skip the knowledge-alignment (configs[3]) leg")
    ap.add_argument("--no-tf32", action="store_true", help="skip the TF32-class arm")
    ap.add_argument("--no-extras", action="store_true", help="skip configs[1] latency and the per-kernel trace")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    from prediff_b200 import _lib as L
    from prediff_b200.diffusion import LatentDiffusion
    from prediff_b200.vae import AutoencoderKL

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the prediff_b200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L.init()
    W = max(args.warmup, 3)
    K, B, S = args.steps, args.batch, args.ddim_steps

    ucfg, vcfg = Wt.UNetConfig(), Wt.VAEConfig()
    unet = build_unet(ucfg, B, "bf16")
    vae = AutoencoderKL(block_out_channels=vcfg.block_out_channels, layers_per_block=vcfg.layers_per_block,
                        latent_channels=vcfg.latent_channels, sample_size=(vcfg.h, vcfg.w), max_frames=B * ucfg.t_in)
    vae.load_state_dict({k: torch.from_numpy(v) for k, v in
                         Wt.seeded_state_dict(Wt.vae_param_spec(vcfg), VAE_SEED).items()}, strict=True)
    ldm = LatentDiffusion(torch_nn_module=unet, first_stage_model=vae, cond_stage_model="__is_first_stage__",
                          data_shape=(ucfg.t_out, vcfg.h, vcfg.w, 1), latent_shape=(ucfg.t_out, ucfg.h, ucfg.w, ucfg.c))

    # ---- kernel-only leg: latents resident in HBM -----------------------------------------------------------
    G = B * world                                                   # global ensemble, sliced by rank
    lo, hi = rank * B, (rank + 1) * B
    zT_all = torch.from_numpy(np_inp(4242, G, ucfg.t_out, ucfg.h, ucfg.w, ucfg.c))
    zT = zT_all[lo:hi].to(dev)
    zc = torch.from_numpy(np_inp(4243, G, ucfg.t_in, ucfg.h, ucfg.w, ucfg.c))[lo:hi].to(dev)
    shape = tuple(zT.shape)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loops(model, n_loops, **kw):
        """n_loops back-to-back device-resident loops, CUDA events on the launching stream, max over ranks (ms)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_loops):
            model.ddim_sample_loop(cond=zc, shape=shape, x_T=zT, ddim_steps=S, eta=0.0, **kw)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for _ in range(W):
        ldm.ddim_sample_loop(cond=zc, shape=shape, x_T=zT, ddim_steps=S, eta=0.0)
    clocks = ClockSampler(local_rank, enabled=(local_rank == 0))
    clocks.start()
    ms = timed_loops(ldm, K)
    value = world * B * S * K / (ms * 1e-3)

    # ---- end-to-end leg: pinned host frames -> sample() -> pinned host forecasts ----------------------------
    # The ensemble's gathered frames live in ONE persistent device buffer: the VAE decoder's last kernel writes this rank's
    # forecasts straight into its slice (sample(out=...)) and the path's single collective - the all-gather of decoded
    # frames - runs in place on it: no allocation, no staging copy.
    y_all = torch.from_numpy(np_inp(780, G, ucfg.t_in, vcfg.h, vcfg.w, 1, uniform=True))
    y_host = y_all[lo:hi].clone().pin_memory()
    zT_host = zT.cpu().pin_memory()
    out_host = torch.empty(G, ucfg.t_out, vcfg.h, vcfg.w, 1).pin_memory()
    gathered = torch.empty(G, ucfg.t_out, vcfg.h, vcfg.w, 1, device=dev)
    mine = gathered[lo:hi]

    def e2e_step():
        y = y_host.to(dev, non_blocking=True)
        zt = zT_host.to(dev, non_blocking=True)
        ldm.sample(cond={"y": y}, batch_size=B, x_T=zt, sampler="ddim", ddim_steps=S, ddim_eta=0.0, out=mine)
        if world > 1:
            dist.all_gather_into_tensor(gathered, mine)
        out_host.copy_(gathered, non_blocking=True)

    for _ in range(W):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        e2e_step()
    e1.record()
    barrier()
    clk = clocks.summary()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = ms_e2e.item()
    e2e_value = world * B * S * K / (ms_e2e * 1e-3)
    h2d = y_host.numel() * 4 + zT_host.numel() * 4
    d2h = out_host.numel() * 4

    # ---- shard invariance: the gathered (G, 6, 128, 128, 1) frames of the N-GPU run
    # equal, bit for bit, what ONE GPU computes for the same global seed (rank 0 recomputes every shard locally) ------
    shard_invariant = None
    if world > 1 and rank == 0:
        ref = torch.empty_like(gathered)
        for r in range(world):
            ldm.sample(cond={"y": y_all[r * B:(r + 1) * B].to(dev)}, batch_size=B, x_T=zT_all[r * B:(r + 1) * B].to(dev),
                       sampler="ddim", ddim_steps=S, ddim_eta=0.0, out=ref[r * B:(r + 1) * B])
        torch.cuda.synchronize()
        shard_invariant = bool(torch.equal(ref, gathered)) and bool(torch.equal(out_host, ref.cpu()))

    # ---- BASELINE.json configs[3]: the same loop with knowledge-alignment guidance (KA forward + backward each step) ----
    ka_line = None
    if not args.no_ka:
        from prediff_b200.alignment import SEVIRAvgIntensityAlignment
        kcfg = Wt.KAConfig()
        al = SEVIRAvgIntensityAlignment(alignment_type="avg_x", guide_scale=kcfg.guide_scale, model_type="cuboid",
                                        model_args=dict(input_shape=[kcfg.t, kcfg.h, kcfg.w, kcfg.c], base_units=kcfg.base_units,
                                                        depth=list(kcfg.depth), block_attn_patterns="axial",
                                                        num_heads=kcfg.num_heads, pool="attention", readout_seq=True,
                                                        out_len=kcfg.t, max_batch=B))
        al.model.load_state_dict({k: torch.from_numpy(v) for k, v in
                                  Wt.seeded_state_dict(Wt.ka_param_spec(kcfg), KA_SEED).items()}, strict=False)
        ldm.set_alignment(al.get_mean_shift)
        kw = {"avg_x_gt": torch.full((B, 1), 0.3, device=dev)}
        Kk = max(2, min(K, 5))
        for _ in range(2):
            ldm.ddim_sample_loop(cond=zc, shape=shape, x_T=zT, ddim_steps=S, eta=0.0, use_alignment=True, alignment_kwargs=kw)
        ms_ka = timed_loops(ldm, Kk, use_alignment=True, alignment_kwargs=kw)
        nf, nb = ctypes.c_int(), ctypes.c_int()
        L.check(L.lib().pd_ka_kernels(al.model.handle, B, ctypes.byref(nf), ctypes.byref(nb)))
        ka_line = {"workload": f"PreDiff-KA: {S}-step DDIM with knowledge-alignment guidance, batch={B} per GPU "
                               "(BASELINE.json configs[3]); KA forward + input-gradient backward every step, run on a "
                               "second stream beside the UNet",
                   "value": world * B * S * Kk / (ms_ka * 1e-3), "unit": UNIT, "loops": Kk,
                   "ms_per_step": ms_ka / Kk, "ka_kernels_per_step": nf.value + nb.value,
                   "gflop_per_sample_step": GFLOP_PER_SAMPLE_STEP + 22.95}
        ldm.set_alignment(None)

    # ---- TF32-class arm: the same loop with kind::tf32 operands (the reference's own GPU arithmetic) --------------
    tf32_line, unet_tf32 = None, None
    if not args.no_tf32:
        unet_tf32 = build_unet(ucfg, B, "tf32")
        ldm_tf32 = LatentDiffusion(torch_nn_module=unet_tf32, latent_shape=(ucfg.t_out, ucfg.h, ucfg.w, ucfg.c))
        Kt = max(2, min(K, 5))
        for _ in range(3):
            ldm_tf32.ddim_sample_loop(cond=zc, shape=shape, x_T=zT, ddim_steps=S, eta=0.0)
        ms_t = timed_loops(ldm_tf32, Kt)
        v_t = world * B * S * Kt / (ms_t * 1e-3)
        tp = load_json("profiles/tf32_peak_r02.json")
        tf32_line = {"value": v_t, "unit": UNIT, "dtype": "tf32", "loops": Kt, "ms_per_step": ms_t / Kt,
                     "numerics": "tcgen05.mma kind::tf32 on fp32 storage rounded to tf32, fp32 accumulate / residual stream / "
                                 "attention core; rel-RMS <= 2e-3 per UNet step and after the 50-step loop vs the fp32 "
                                 "reference (tests/test_tf32_gpu.py) - like-for-like with the reference's TF32 runs",
                     "roofline": None if tp is None else {
                         "bound": "tensor", "achieved": v_t / world * GFLOP_PER_SAMPLE_STEP * 1e-3,
                         "peak": tp["tf32_tflops_sustained"], "unit": "TFLOP/s",
                         "frac": v_t / world * GFLOP_PER_SAMPLE_STEP * 1e-3 / tp["tf32_tflops_sustained"],
                         "peak_source": "measured sustained cuBLAS TF32 8192^3 on this pool (profiles/tf32_peak_r02.json; "
                                        f"burst {tp['tf32_tflops']:.0f})"}}

    # ---- per-launch device time of one UNet forward + BASELINE.json configs[1] -----------------------------------
    nk = ctypes.c_int()
    n_sub = L.lib().pd_sampler_sub_batches(ldm._sampler, B)   # concurrent sub-batches inside the loop
    L.check(L.lib().pd_unet_kernels_per_forward(unet.handle, B // n_sub, ctypes.byref(nk)))
    launches_per_loop_step = n_sub * (nk.value + 1) + 1   # per sub-batch: UNet + sampler_update; + advance_step
    gpu_launches = launches_per_loop_step * S * K
    extras = {}
    if not args.no_extras:
        t = torch.full((B,), 981, device=dev, dtype=torch.int64)
        eps = torch.empty_like(zT)
        # a CUDA-graph replay of one forward with a %globaltimer stamp after every launch (tools/trace_unet.py --graph):
        # launches are issued by the GPU front end, so kernels shorter than the ~5 us host launch cost are timed correctly
        trace, slot = graph_trace(unet, B, zT, t, zc, eps)
        extras["trace"] = (trace, slot)
        if rank == 0 and world == 1:
            extras["b1"] = single_step_latency(unet, ucfg, dev)
            # the same step on a model built for single samples: 72 stream-K CTAs per sample instead of the batch-invariant 36
            # (pd_unet_set_streamk_ctas; another fixed summation order, so not bit-identical to the batch-4 cut)
            unet_b1 = build_unet(ucfg, 1, "bf16", streamk_ctas_per_sample=72)
            extras["b1"]["latency_cut"] = {k: v for k, v in single_step_latency(unet_b1, ucfg, dev).items()
                                           if k in ("ms", "value", "tflops")}
            extras["b1"]["latency_cut"]["streamk_ctas_per_sample"] = 72
            del unet_b1
            if unet_tf32 is not None:
                extras["b1_tf32"] = single_step_latency(unet_tf32, ucfg, dev, iters=100)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    achieved = value / world * GFLOP_PER_SAMPLE_STEP * 1e-3            # TFLOP/s per GPU, algorithmic FLOPs
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic latents / frames (numpy PCG64 seeds), seeded random weights of the shipped SEVIR-LR architecture",
        "config": {"workload": f"SEVIR-LR 50-step DDIM (eta=0) p_sample_loop, batch={B} per GPU (BASELINE.json configs[2])",
                   "step": f"one {S}-step DDIM loop over a batch of {B} forecasts per GPU = {B * S} UNet evaluations",
                   "ensemble": f"{G} members, {B} per rank, no data-path collective; one in-place all-gather of decoded frames in e2e",
                   "l2": "inputs exceed L2: 274 MB of bf16 UNet weights are re-streamed every denoise step (L2 = 126 MB)",
                   "numerics": "bf16 tensor-core operands, fp32 accumulate, fp32 residual stream (rel-RMS 7e-3 per step vs the "
                               "fp32 reference); the like-for-like TF32 arm is in tf32_arm",
                   "cuda_graph": f"the whole {S}-step loop is one graph launch", "sub_batches": n_sub},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / K, "api": "LatentDiffusion.sample(cond={'y': frames}, sampler='ddim', out=<slice of the "
                "gathered buffer>) + in-place all_gather_into_tensor"},
        "gpu_launches": gpu_launches,
        "clocks": clk,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tensor_tflops"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tensor_tflops"], "traffic": None,
                     "peak_source": f"{peaks['source']} sustained bf16 (MEASURED_PEAKS.json)",
                     "definition": "sample-steps/s per GPU x 653.43 GFLOP (SURVEY.md 8d) / sustained bf16 peak"},
    }
    if shard_invariant is not None:
        line["shard_invariant"] = shard_invariant
    if "trace" in extras:
        trace, slot = extras["trace"]
        gemm_sites = ("conv1", "conv2", "qkv", "proj", "ffn1", "ffn2", "proj_ffn_fused", "ffn_fused", "proj_ffn_cluster",
                      "ffn_cluster", "qkv_attn_T", "qkv_attn_H", "qkv_attn_W", "skip", "up.conv", "down.reduction", "final.proj")
        is_gemm = lambda k: k.split(".")[-1] in gemm_sites or k in gemm_sites   # noqa: E731
        n_launch = sum(v[0] for v in trace.values())
        kern_ms = sum(v[1] for v in trace.values()) * 1e-3
        step_ms = ms / K / S                                  # one denoise step inside the timed loop
        # what the trace's slot subtraction hides: per-node dispatch overhead inside the loop's graph
        overhead_ms = max(step_ms - kern_ms, 0.0)
        per_launch_overhead_us = 1e3 * overhead_ms / max(n_launch, 1)
        dram = load_json("profiles/kernel_dram_r02.json") or {}
        kernels = []
        for name, (cnt, us, fl) in sorted(trace.items(), key=lambda kv: -kv[1][1]):
            if cnt == 0 or us <= 0:
                continue
            ent = {"site": name, "launches": cnt, "avg_us": us / cnt, "share_of_step": us * 1e-3 / step_ms}
            if fl > 0:
                ent["gflop_per_launch"] = fl / cnt * 1e-9
                ent["tflops"] = fl / (us * 1e-6) * 1e-12
                ent["frac_of_peak"] = ent["tflops"] / peaks["tensor_tflops"]
                ent["frac_incl_launch_overhead"] = fl / ((us + cnt * per_launch_overhead_us) * 1e-6) * 1e-12 / peaks["tensor_tflops"]
            ent["algorithmic_bytes_per_launch"] = algorithmic_bytes(name, B)
            if name in dram:
                ent["dram_bytes_per_launch"] = dram[name].get("dram_bytes")   # ncu pass at batch 4 (profiles/kernel_dram_r02.json)
                ent["kernel"] = dram[name].get("kernel")
            kernels.append(ent)
        g_us = sum(v[1] for k, v in trace.items() if is_gemm(k))
        g_fl = sum(v[2] for k, v in trace.items() if is_gemm(k))
        g_n = sum(v[0] for k, v in trace.items() if is_gemm(k))
        stack_ms = sum(v[1] for k, v in trace.items() if ".stack." in k) * 1e-3
        top = kernels[0] if kernels else None
        line["roofline"].update({
            "traffic": top.get("dram_bytes_per_launch") if top else None,
            "traffic_kernel": top["site"] if top else None,
            "kernel_timing": f"per-launch %globaltimer stamps inside a CUDA-graph replay of one batch-{B} forward; avg_us = "
                             f"stamp interval - the stamp kernel's own slot ({slot:.2f} us); launch overhead reported separately",
            "launch_overhead": {"launches_per_step": n_launch, "kernel_ms_per_step": kern_ms, "in_loop_step_ms": step_ms,
                                "overhead_ms_per_step": overhead_ms, "per_launch_us": per_launch_overhead_us,
                                "share_of_step": overhead_ms / step_ms},
            "dominant_kernel": {"name": "tcgen05 implicit-GEMM family: conv_streamk_kernel, ffn_fused_kernel, ffn_cluster_kernel, "
                                        "qkv_attn_kernel, gemm_tc_kernel (conv3d / conv2d / linear / fused transformer layers)",
                                "launches_per_forward": int(g_n), "ms_per_forward": g_us * 1e-3,
                                "avg_launch_us": g_us / max(g_n, 1), "tflops": g_fl / (g_us * 1e-6) * 1e-12,
                                "frac_of_peak": g_fl / (g_us * 1e-6) * 1e-12 / peaks["tensor_tflops"],
                                "frac_incl_launch_overhead": g_fl / ((g_us + g_n * per_launch_overhead_us) * 1e-6) * 1e-12
                                / peaks["tensor_tflops"],
                                "share_of_step": g_us * 1e-3 / step_ms},
            "kernels": kernels[:16],
            # north_star's "cuboid-attention roofline": the StackCuboidSelfAttentionBlock subset (LayerNorm, QKV, axial
            # attention, projection, FFN of all 24 + 24 layers = 252.9 of the 653.43 GFLOP), in place
            "cuboid_attention_blocks": {
                "gflop_per_sample_step": CUBOID_GFLOP_PER_SAMPLE_STEP, "ms_per_forward": stack_ms,
                "tflops": CUBOID_GFLOP_PER_SAMPLE_STEP * B / stack_ms,
                "frac_of_peak": CUBOID_GFLOP_PER_SAMPLE_STEP * B / stack_ms / peaks["tensor_tflops"],
                "share_of_step": stack_ms / step_ms}})
    if "b1" in extras:
        b1 = extras["b1"]
        b1["frac_of_peak"] = b1["tflops"] / peaks["tensor_tflops"]
        if "b1_tf32" in extras:
            b1["tf32"] = {k: extras["b1_tf32"][k] for k in ("ms", "value", "tflops", "iters")}
        line["single_step_b1"] = b1
    if tf32_line is not None:
        tf32_line["vs_bf16_arm"] = tf32_line["value"] / value
        line["tf32_arm"] = tf32_line
    if ka_line is not None:
        ka_line["slowdown_vs_unguided"] = (ka_line["ms_per_step"]) / (ms / K)
        line["knowledge_alignment"] = ka_line
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(args)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
