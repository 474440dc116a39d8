#!/usr/bin/env python
"""Print the roofline-relevant metrics of every kernel in an `ncu --page raw --csv` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64pipe%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
        ("smsp__inst_executed.sum", "warp_inst"),
        ("l1tex__t_sector_hit_rate.pct", "L1hit%"), ("lts__t_sector_hit_rate.pct", "L2hit%"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sectors"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ]
idx = {w: hdr.index(w) for w, _ in want if w in hdr}
ki = hdr.index("Kernel Name")
seen = set()
for d in data:
    name = d[ki].split("(")[0]
    if name in seen and "--all" not in sys.argv:
        continue
    seen.add(name)
    print("---", name[:110])
    print("   " + "  ".join(f"{lbl}={d[idx[w]]}{units[idx[w]] if units[idx[w]] not in ('%','') else ''}" for w, lbl in want if w in idx))
