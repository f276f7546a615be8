"""Pivots an `ncu --metrics ... --csv` log (one row per launch x metric) into one line per launch."""
import csv
import re
import sys
from collections import OrderedDict

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
L = OrderedDict()
for x in csv.DictReader(rows):
    d = L.setdefault(x["ID"], {"name": re.sub(r"\(.*", "", x["Kernel Name"]).split("::")[-1]})
    d[x["Metric Name"]] = x["Metric Value"].replace(",", "")


def f(d, k, scale=1.0):
    try:
        return float(d.get(k, "nan")) * scale
    except ValueError:
        return float("nan")


print(f"{'#':>3s} {'kernel':38s} {'grid':>5s} {'blk':>4s} {'regs':>4s} {'smemKB':>6s} {'us':>7s} {'dramR MB':>8s} {'dramW MB':>8s} "
      f"{'L2 MB':>7s} {'tensor%':>7s} {'dram%':>6s} {'warps%':>6s} {'smem wavefronts':>15s}")
for k, d in L.items():
    print(f"{k:>3s} {d['name'][:38]:38s} {f(d, 'launch__grid_size'):5.0f} {f(d, 'launch__block_size'):4.0f} "
          f"{f(d, 'launch__registers_per_thread'):4.0f} {f(d, 'launch__shared_mem_per_block_dynamic', 1 / 1024):6.0f} "
          f"{f(d, 'gpu__time_duration.sum', 1e-3):7.1f} {f(d, 'dram__bytes_read.sum', 1e-6):8.2f} "
          f"{f(d, 'dram__bytes_write.sum', 1e-6):8.2f} {f(d, 'lts__t_bytes.sum', 1e-6):7.1f} "
          f"{f(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):7.1f} "
          f"{f(d, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{f(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{f(d, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'):15.0f}")
