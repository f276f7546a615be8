"""Turns an `ncu --metrics ... --csv` pass over one eager UNet forward (tools/profile_unet.py) plus the plan labels of the
same forward (tools/dump_unet_labels.py) into per-call-site rows: kernel, launches, device time, DRAM bytes per launch,
L2 bytes, tensor-pipe %, algorithmic FLOPs -> profiles/ncu_<tag>_unet_sites.txt and profiles/kernel_dram_<tag>.json (read
by bench.py for the per-kernel roofline entries). For a VAE pass (no labels) rows are grouped by kernel name + grid.

  python tools/ncu_sites.py unet gpurun_out/ncu_r02_unet_fwd_b4.csv gpurun_out/unet_labels_r02_b4.txt r02
  python tools/ncu_sites.py vae  gpurun_out/ncu_r02_vae_b4.csv r02
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6550.0


def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    out = collections.OrderedDict()
    for x in csv.DictReader(rows):
        d = out.setdefault(x["ID"], {"name": re.sub(r"\(.*", "", x["Kernel Name"]).split("::")[-1]})
        try:
            d[x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
        except ValueError:
            d[x["Metric Name"]] = float("nan")
    return list(out.values())


def g(d, k, scale=1.0):
    return d.get(k, float("nan")) * scale


TEN = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"

mode = sys.argv[1]
if mode == "unet":
    L = launches(sys.argv[2])
    labs = [l.split() for l in open(sys.argv[3]) if l.strip()]
    labs = [(a, float(b)) for a, b in labs if a not in ("memset",)]   # the memset step is not a kernel
    tag = sys.argv[4]
    assert len(L) == len(labs), (len(L), len(labs))
    agg = collections.OrderedDict()
    for d, (site, fl) in zip(L, labs):
        a = agg.setdefault(site, {"kernel": d["name"], "n": 0, "us": 0.0, "dram": 0.0, "l2": 0.0, "tensor": 0.0, "flops": 0.0,
                                  "grid": int(g(d, "launch__grid_size"))})
        a["n"] += 1
        a["us"] += g(d, "gpu__time_duration.sum", 1e-3)
        a["dram"] += g(d, "dram__bytes_read.sum") + g(d, "dram__bytes_write.sum")
        a["l2"] += g(d, "lts__t_bytes.sum")
        a["tensor"] += g(d, TEN)
        a["flops"] += fl
    tot = sum(a["us"] for a in agg.values())
    lines = [f"ncu metric pass over one eager UNet forward, batch 4, bf16 operands ({len(L)} launches, {tot:.0f} us serialised; cold-cache, "
             "serialised replays: compare SHARES with the graph trace, not absolutes)",
             f"{'site':26s} {'kernel':30s} {'n':>3s} {'grid':>5s} {'avg us':>7s} {'share':>6s} {'dram MB/launch':>14s} {'L2 MB/launch':>12s} "
             f"{'tensor%':>7s} {'GFLOP/launch':>12s} {'dram GB/s':>9s}"]
    js = {}
    for site, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = a["n"]
        lines.append(f"{site:26s} {a['kernel'][:30]:30s} {n:3d} {a['grid']:5d} {a['us'] / n:7.1f} {100 * a['us'] / tot:5.1f}% "
                     f"{a['dram'] / n * 1e-6:14.2f} {a['l2'] / n * 1e-6:12.1f} {a['tensor'] / n:7.1f} {a['flops'] / n * 1e-9:12.2f} "
                     f"{a['dram'] / (a['us'] * 1e-6) * 1e-9:9.0f}")
        js[site] = {"kernel": a["kernel"], "dram_bytes": a["dram"] / n, "l2_bytes": a["l2"] / n,
                    "tensor_pipe_pct": a["tensor"] / n, "ncu_us": a["us"] / n}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", f"ncu_{tag}_unet_sites.txt"), "w").write("\n".join(lines) + "\n")
    json.dump(js, open(os.path.join(ROOT, "profiles", f"kernel_dram_{tag}.json"), "w"), indent=1)
    print("\n".join(lines[:24]))
else:
    L = launches(sys.argv[2])
    tag = sys.argv[3]
    agg = collections.OrderedDict()
    for d in L:
        key = (d["name"], int(g(d, "launch__grid_size")))
        a = agg.setdefault(key, {"n": 0, "us": 0.0, "dr": 0.0, "dw": 0.0, "l2": 0.0, "tensor": 0.0})
        a["n"] += 1
        a["us"] += g(d, "gpu__time_duration.sum", 1e-3)
        a["dr"] += g(d, "dram__bytes_read.sum")
        a["dw"] += g(d, "dram__bytes_write.sum")
        a["l2"] += g(d, "lts__t_bytes.sum")
        a["tensor"] += g(d, TEN)
    tot = sum(a["us"] for a in agg.values())
    lines = [f"ncu metric pass over AutoencoderKL.encode (28 frames) + decode (24 frames), shipped sizes ({len(L)} launches, {tot:.0f} us "
             f"serialised). HBM peak = {HBM:.0f} GB/s (MEASURED_PEAKS.json).",
             f"{'kernel':34s} {'grid':>6s} {'n':>3s} {'avg us':>8s} {'share':>6s} {'dramR MB':>9s} {'dramW MB':>9s} {'GB/s':>6s} {'of HBM':>6s} "
             f"{'L2 MB':>8s} {'tensor%':>7s}"]
    by = collections.defaultdict(lambda: [0.0, 0.0])
    for (name, grid), a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = a["n"]
        gbs = (a["dr"] + a["dw"]) / (a["us"] * 1e-6) * 1e-9
        lines.append(f"{name[:34]:34s} {grid:6d} {n:3d} {a['us'] / n:8.1f} {100 * a['us'] / tot:5.1f}% {a['dr'] / n * 1e-6:9.2f} "
                     f"{a['dw'] / n * 1e-6:9.2f} {gbs:6.0f} {gbs / HBM:6.2f} {a['l2'] / n * 1e-6:8.1f} {a['tensor'] / n:7.1f}")
        by[name][0] += a["us"]
        by[name][1] += a["dr"] + a["dw"]
    lines.append("")
    lines.append("per kernel (all grids): total us, share, achieved DRAM GB/s, fraction of the measured HBM peak")
    for name, (us, byts) in sorted(by.items(), key=lambda kv: -kv[1][0]):
        gbs = byts / (us * 1e-6) * 1e-9
        lines.append(f"{name[:34]:34s} {us:9.1f} {100 * us / tot:5.1f}% {gbs:7.0f} GB/s {gbs / HBM:5.2f}")
    open(os.path.join(ROOT, "profiles", f"ncu_{tag}_vae_summary.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
