#!/usr/bin/env python
"""32-byte record gather rates of this GPU (shamb200_microbench 10+p / 20+p): the roof of the L1 data pipe for
the access pattern of the SPH neighbour loops.  One JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shamrock_b200 import _capi  # noqa: E402

NAMES = ["coalesced", "random_record_per_lane", "random_line_bankgroup_eq_lane", "random_line_one_bankgroup",
         "four_lanes_per_line"]
ctx = _capi.Context(0)
out = {"unit": "G records (32 B) / s, whole GPU"}
for base, where in ((10, "L1"), (20, "L2_32MiB")):
    out[where] = {n: round(ctx.microbench(base + p), 1) for p, n in enumerate(NAMES)}
out["fp64_tflops"] = round(ctx.microbench("fp64"), 2)
print(json.dumps(out))
