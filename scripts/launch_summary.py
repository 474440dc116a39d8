#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of the last step."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [x["Kernel Name"] for x in rows]
idx = [i for i, n in enumerate(names) if "predictor_kernel" in n]
last = rows[idx[-1]:] if idx else rows
agg = collections.OrderedDict()
for x in last:
    n = re.sub(r"\(.*", "", x["Kernel Name"])
    t = float(x["Metric Value"].replace(",", ""))
    u = x["Metric Unit"]
    t = t / 1e6 if u == "ns" else (t / 1e3 if u == "us" else t)
    agg.setdefault(n, [0, 0]); agg[n][0] += t; agg[n][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"last step: {len(last)} launches, {tot:.3f} ms of kernel time (cold-cache, serialised)")
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{v[0]:9.3f} ms {100*v[0]/tot:5.1f}% x{v[1]:3d} {n[:100]}")
