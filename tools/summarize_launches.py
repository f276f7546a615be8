"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per (kernel, grid) totals + sequence."""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0])
seq = []
for row in rows:
    name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1]
    us = float(row["Metric Value"].replace(",", "")) / 1000.0
    agg[(name, row["Grid Size"])][0] += 1
    agg[(name, row["Grid Size"])][1] += us
    seq.append((name, row["Grid Size"], us))
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {len(seq)} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v[1]:9.1f} us {100 * v[1] / tot:5.1f}%  n={v[0]:3d} avg={v[1] / v[0]:7.1f}  {k[0][:44]:44s} grid={k[1]}")
if len(sys.argv) > 3:
    a, b = map(int, sys.argv[3].split(":"))
    for i, (n, g, us) in enumerate(seq[a:b]):
        print(i + a, n[:40], g, f"{us:.1f}")
