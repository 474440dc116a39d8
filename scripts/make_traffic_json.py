#!/usr/bin/env python
"""profiles/traffic_17M.json from an `ncu --set full --page raw --csv` export of the heavy kernels of one step at
the bench size: DRAM bytes per launch, kernel time and the utilisation of the units that bind each kernel.
bench.py attaches these numbers to its stages only while the capture's kernel time agrees with the live stage
time within 5 % (a capture of other code is not evidence).

    python scripts/make_traffic_json.py gpurun_out/ncu_raw_TAG.csv NPART > profiles/traffic_17M.json
"""
import csv
import json
import sys

STAGE_OF = {"group_walk_kernel": "neigh_walk", "neigh_lists_kernel": "neigh_lists", "h_solve_fast_kernel": "h_iteration",
            "av_operators_fast_kernel": "divv_curlv_dtdivv", "force_cfl_fast_kernel": "forces"}
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")


def val(d, name, scale=1.0):
    i = hdr.index(name)
    u = units[i]
    f = float(d[i].replace(",", ""))
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return f * mult * scale


out = {"npart": int(sys.argv[2]),
       "source": f"ncu --set full --clock-control none, one launch of each kernel in the 4th step of `bench.py "
                 f"--npart-per-gpu 16777216` ({sys.argv[1]}); dram__bytes_read.sum + dram__bytes_write.sum per launch; "
                 "l1_lsu_data_pipe_pct = l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
       "stages": {}}
for d in data:
    name = d[ki].split("(")[0]
    for k, stage in STAGE_OF.items():
        if k in name and stage not in out["stages"]:
            rd, wr = val(d, "dram__bytes_read.sum"), val(d, "dram__bytes_write.sum")
            out["stages"][stage] = {
                "kernel": k, "dram_bytes_read": rd, "dram_bytes_write": wr, "ncu_time_ms": val(d, "gpu__time_duration.sum"),
                "traffic": rd + wr,
                "fp64_pipe_pct": val(d, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                "l1_lsu_data_pipe_pct": val(d, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                "issue_active_pct": val(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "registers": val(d, "launch__registers_per_thread"),
                "lanes_per_inst": val(d, "smsp__thread_inst_executed_per_inst_executed.ratio")}
# the "neigh_walk" stage of bench.py is the group walk followed by the Euclidean cull of its candidates
for d in data:
    if "euclid_cull_kernel" in d[ki] and "neigh_walk" in out["stages"] and "cull_ms" not in out["stages"]["neigh_walk"]:
        st = out["stages"]["neigh_walk"]
        st["cull_ms"] = val(d, "gpu__time_duration.sum")
        st["walk_ms"] = st["ncu_time_ms"]
        st["ncu_time_ms"] = st["walk_ms"] + st["cull_ms"]
        st["traffic"] += val(d, "dram__bytes_read.sum") + val(d, "dram__bytes_write.sum")
        st["kernel"] = "group_walk_kernel + euclid_cull_kernel (unit utilisations: group_walk_kernel)"
print(json.dumps(out, indent=1))
