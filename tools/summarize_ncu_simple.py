"""Per-launch table from an `ncu --metrics ... --csv` log: id, kernel, duration, grid, registers, warps active, instructions.
  python tools/summarize_ncu_simple.py gpurun_out/ncu_gv.csv"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
d = OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii], r[ki].split("(")[0].replace("void ", "").replace("pd::<unnamed>::", "")[:48]), {})[r[mi]] = r[vi]
print(f"{'id':>3} {'kernel':48} {'us':>8} {'grid':>6} {'regs':>5} {'warps%':>7} {'dram MB':>8}")
for (i, k), m in d.items():
    dram = (float(m.get("dram__bytes_read.sum", 0) or 0) + float(m.get("dram__bytes_write.sum", 0) or 0)) / 1e6
    print(f"{i:>3} {k:48} {float(m.get('gpu__time_duration.sum', 0)) / 1e3:8.2f} {m.get('launch__grid_size', ''):>6} "
          f"{m.get('launch__registers_per_thread', ''):>5} {float(m.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0) or 0):7.1f} {dram:8.2f}")
