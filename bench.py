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
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=3)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


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
    sample = f"{args.steps} of the {args.ddim_steps} denoise steps of one batch-{B} DDIM loop (UNet forward + DDIM update)"
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


def graph_trace(unet, B, x, t, cond, out):
    """{site label: (launches, us)} of one UNet forward replayed as a CUDA graph with a stamp kernel after every launch."""
    import collections
    import numpy as np
    import torch
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
    d = np.diff(ns.cpu().numpy()[:n[0] + 1]).astype(np.float64) * 1e-3
    slot = float(d.min())
    agg = collections.OrderedDict()
    for name, us in zip(lab, d):
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us - slot
    return agg


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
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    from prediff_b200 import _lib as L
    from prediff_b200.diffusion import LatentDiffusion
    from prediff_b200.dist import sample_ensemble
    from prediff_b200.unet import CuboidTransformerUNet
    from prediff_b200.vae import AutoencoderKL
    import ctypes

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
    unet = CuboidTransformerUNet([ucfg.t_in, ucfg.h, ucfg.w, ucfg.c], [ucfg.t_out, ucfg.h, ucfg.w, ucfg.c],
                                 base_units=ucfg.base_units, depth=list(ucfg.depth), num_heads=ucfg.num_heads,
                                 block_attn_patterns="axial", max_batch=B)
    unet.load_state_dict({k: torch.from_numpy(v) for k, v in
                          Wt.seeded_state_dict(Wt.unet_param_spec(ucfg), UNET_SEED).items()}, strict=False)
    vae = AutoencoderKL(block_out_channels=vcfg.block_out_channels, layers_per_block=vcfg.layers_per_block,
                        latent_channels=vcfg.latent_channels, sample_size=(vcfg.h, vcfg.w), max_frames=B * ucfg.t_in)
    vae.load_state_dict({k: torch.from_numpy(v) for k, v in
                         Wt.seeded_state_dict(Wt.vae_param_spec(vcfg), VAE_SEED).items()}, strict=True)
    ldm = LatentDiffusion(torch_nn_module=unet, first_stage_model=vae, cond_stage_model="__is_first_stage__",
                          data_shape=(ucfg.t_out, vcfg.h, vcfg.w, 1), latent_shape=(ucfg.t_out, ucfg.h, ucfg.w, ucfg.c))

    # ---- kernel-only leg: latents resident in HBM -----------------------------------------------------------
    G = B * world                                                   # global ensemble, sliced by rank
    zT = torch.from_numpy(np_inp(4242, G, ucfg.t_out, ucfg.h, ucfg.w, ucfg.c))[rank * B:(rank + 1) * B].to(dev)
    zc = torch.from_numpy(np_inp(4243, G, ucfg.t_in, ucfg.h, ucfg.w, ucfg.c))[rank * B:(rank + 1) * B].to(dev)
    shape = tuple(zT.shape)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        ldm.ddim_sample_loop(cond=zc, shape=shape, x_T=zT, ddim_steps=S, eta=0.0)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        z0 = ldm.ddim_sample_loop(cond=zc, shape=shape, x_T=zT, ddim_steps=S, eta=0.0)
    e1.record()
    barrier()
    clk = clocks.summary()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    value = world * B * S * K / (ms * 1e-3)

    # ---- end-to-end leg: pinned host frames -> sample() -> pinned host forecasts ----------------------------
    y_host = torch.from_numpy(np_inp(780, G, ucfg.t_in, vcfg.h, vcfg.w, 1, uniform=True))[rank * B:(rank + 1) * B].pin_memory()
    zT_host = zT.cpu().pin_memory()
    out_host = torch.empty(G, ucfg.t_out, vcfg.h, vcfg.w, 1).pin_memory()

    def e2e_step():
        y = y_host.to(dev, non_blocking=True)
        zt = zT_host.to(dev, non_blocking=True)
        frames = ldm.sample(cond={"y": y}, batch_size=B, x_T=zt, sampler="ddim", ddim_steps=S, ddim_eta=0.0)
        if world > 1:   # the path's one collective: all-gather of the decoded frames
            full = torch.empty((G,) + tuple(frames.shape[1:]), device=dev)
            dist.all_gather_into_tensor(full, frames.contiguous())
        else:
            full = frames
        out_host.copy_(full, non_blocking=True)
        return full

    for _ in range(W):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    ms_e2e = ms_e2e.item()
    e2e_value = world * B * S * K / (ms_e2e * 1e-3)
    h2d = y_host.numel() * 4 + zT_host.numel() * 4
    d2h = out_host.numel() * 4

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
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(Kk):
            ldm.ddim_sample_loop(cond=zc, shape=shape, x_T=zT, ddim_steps=S, eta=0.0, use_alignment=True, alignment_kwargs=kw)
        e1.record()
        barrier()
        ms_ka = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms_ka, op=dist.ReduceOp.MAX)
        nf, nb = ctypes.c_int(), ctypes.c_int()
        L.check(L.lib().pd_ka_kernels(al.model.handle, B, ctypes.byref(nf), ctypes.byref(nb)))
        ka_line = {"workload": f"PreDiff-KA: {S}-step DDIM with knowledge-alignment guidance, batch={B} per GPU "
                               "(BASELINE.json configs[3]); KA forward + input-gradient backward every step, run on a "
                               "second stream beside the UNet",
                   "value": world * B * S * Kk / (ms_ka.item() * 1e-3), "unit": UNIT, "loops": Kk,
                   "ms_per_step": ms_ka.item() / Kk, "ka_kernels_per_step": nf.value + nb.value,
                   "gflop_per_sample_step": GFLOP_PER_SAMPLE_STEP + 22.95}
        ldm.set_alignment(None)

    # ---- per-kernel-class device time of one UNet forward (events around every launch) -----------------------
    stats = (ctypes.c_double * 5)()
    t = torch.full((B,), 981, device=dev, dtype=torch.int64)
    eps = torch.empty_like(zT)
    for _ in range(2):
        L.check(L.lib().pd_unet_profile_forward(unet.handle, L.ptr(zT), L.ptr(t), L.ptr(zc), L.ptr(eps), B,
                                                L.stream_ptr(), stats))
    gemm_ms, n_gemm, other_ms, n_other, gemm_flops = list(stats)
    # the same forward as a CUDA-graph replay with a %globaltimer stamp after every launch (tools/trace_unet.py --graph):
    # launches are issued by the GPU front end, so kernels shorter than the ~5 us host launch cost are timed correctly
    trace = graph_trace(unet, B, zT, t, zc, eps)
    gemm_sites = ("conv1", "conv2", "qkv", "proj", "ffn1", "ffn2", "proj_ffn_fused", "ffn_fused", "skip", "up.conv",
                  "down.reduction", "final.proj")
    tr_gemm = [v for k, v in trace.items() if k.split(".")[-1] in gemm_sites or k in gemm_sites]
    tr_other = [v for k, v in trace.items() if not (k.split(".")[-1] in gemm_sites or k in gemm_sites)]
    tr_stack = [v for k, v in trace.items() if ".stack." in k]
    gemm_ms = sum(v[1] for v in tr_gemm) * 1e-3
    other_ms = sum(v[1] for v in tr_other) * 1e-3
    n_gemm, n_other = sum(v[0] for v in tr_gemm), sum(v[0] for v in tr_other if v[1] > 0)
    stack_ms = sum(v[1] for v in tr_stack) * 1e-3
    nk = ctypes.c_int()
    n_sub = L.lib().pd_sampler_sub_batches(ldm._sampler, B)   # concurrent sub-batches inside the loop
    L.check(L.lib().pd_unet_kernels_per_forward(unet.handle, B // n_sub, ctypes.byref(nk)))
    launches_per_loop_step = n_sub * (nk.value + 1) + 1   # per sub-batch: UNet + sampler_update; + advance_step
    gpu_launches = launches_per_loop_step * S * K

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    achieved = value / world * GFLOP_PER_SAMPLE_STEP * 1e-3            # TFLOP/s per GPU, algorithmic FLOPs
    traffic = None
    prof_json = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(prof_json):
        traffic = json.load(open(prof_json)).get("gemm_tc_kernel_dram_bytes_per_launch")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic latents / frames (numpy PCG64 seeds), seeded random weights of the shipped SEVIR-LR architecture",
        "config": {"workload": f"SEVIR-LR 50-step DDIM (eta=0) p_sample_loop, batch={B} per GPU (BASELINE.json configs[2])",
                   "step": f"one {S}-step DDIM loop over a batch of {B} forecasts per GPU = {B * S} UNet evaluations",
                   "ensemble": f"{G} members, {B} per rank, no data-path collective; one all-gather of decoded frames in e2e",
                   "l2": "inputs exceed L2: 274 MB of bf16 UNet weights are re-streamed every denoise step (L2 = 126 MB)",
                   "numerics": "bf16 tensor-core operands, fp32 accumulate, fp32 residual stream", "cuda_graph": True,
                   "sub_batches": n_sub},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / K, "api": "LatentDiffusion.sample(cond={'y': frames}, sampler='ddim')"},
        "gpu_launches": gpu_launches,
        "clocks": clk,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tensor_tflops"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tensor_tflops"], "traffic": traffic,
                     "peak_source": f"{peaks['source']} sustained bf16 (MEASURED_PEAKS.json)",
                     "definition": "sample-steps/s per GPU x 653.43 GFLOP (SURVEY.md 8d) / sustained bf16 peak",
                     "dominant_kernel": {"name": "tcgen05 implicit-GEMM family: gemm_tc_kernel, gemm_tc_persistent_kernel, conv_streamk_kernel, ffn_fused_kernel (conv3d/conv2d/linear)",
                                         "launches_per_forward": int(n_gemm), "ms_per_forward": gemm_ms,
                                         "avg_launch_us": 1e3 * gemm_ms / max(n_gemm, 1),
                                         "tflops": gemm_flops / (gemm_ms * 1e-3) * 1e-12,
                                         "frac_of_peak": gemm_flops / (gemm_ms * 1e-3) * 1e-12 / peaks["tensor_tflops"],
                                         "share_of_forward": gemm_ms / (gemm_ms + other_ms)},
                     "other_kernels": {"launches_per_forward": int(n_other), "ms_per_forward": other_ms},
                     "kernel_timing": f"per-launch globaltimer stamps inside a CUDA-graph replay of one batch-{B} forward "
                                      "(one stamp-kernel slot subtracted per launch)",
                     # north_star's "cuboid-attention roofline": the StackCuboidSelfAttentionBlock subset (LayerNorm, QKV,
                     # axial attention, projection, FFN of all 24 + 24 layers = 252.9 of the 653.43 GFLOP),
                     # timed in place inside the same replay
                     "cuboid_attention_blocks": {
                         "gflop_per_sample_step": CUBOID_GFLOP_PER_SAMPLE_STEP, "ms_per_forward": stack_ms,
                         "tflops": CUBOID_GFLOP_PER_SAMPLE_STEP * B / stack_ms,   # GFLOP / ms = TFLOP/s
                         "frac_of_peak": CUBOID_GFLOP_PER_SAMPLE_STEP * B / stack_ms / peaks["tensor_tflops"],
                         "share_of_forward": stack_ms / (gemm_ms + other_ms)}},
    }
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
