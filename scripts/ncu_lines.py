#!/usr/bin/env python
"""Per-CUDA-source-line instruction and stall-sample shares from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name K > f.csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
out, hdr = [], None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    ii, sa = hdr.index("Instructions Executed"), hdr.index("# Samples")
    try:
        out.append((int(r[0]), r[1], float(r[ii]), float(r[sa])))
    except ValueError:
        pass
ti, ts = sum(o[2] for o in out), sum(o[3] for o in out)
print(f"total warp-inst {ti:.4g}, samples {ts:.0f}")
for ln, src, ni, ns in out:
    if ni > thr * ti or ns > thr * ts:
        print(f"{100*ni/ti:5.1f}% inst {100*ns/ts:5.1f}% smp  L{ln:<4d} {src.strip()[:100]}")
