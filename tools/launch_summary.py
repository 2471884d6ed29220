"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share.
usage: python tools/launch_summary.py launches.csv [last_n_launches]"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        rows.append((r["Kernel Name"], ns, r.get("Grid Size", ""), r.get("Block Size", "")))
if len(sys.argv) > 2:
    rows = rows[-int(sys.argv[2]):]
agg = OrderedDict()
for name, ns, g, bl in rows:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    a = agg.setdefault(short, [0, 0.0])
    a[0] += 1; a[1] += ns
tot = sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.1f} us total (serialised, cold-cache under ncu: compare shares, not absolutes)")
for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ns / 1e3:10.1f} us  {100 * ns / tot:5.1f}%  x{c:<4d} {k[:110]}")
