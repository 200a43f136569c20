#!/usr/bin/env python
"""Key metrics of an .ncu-rep (first instance of every distinct kernel in the report): time, DRAM/L2 traffic, occupancy,
SIMD efficiency, issue utilisation and the warp-stall breakdown. Usage: ncu_summary.py rep [rep...]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "l2_rd_sectors"),
    ("lts__t_sectors_srcunit_tex_op_write.sum", "l2_wr_sectors"),
    ("lts__t_sector_hit_rate.pct", "l2_hit%"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit%"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall_not_selected"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall_branch"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall_no_inst"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall_dispatch"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall_membar"),
]


def summarize(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res, seen = [], set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if name in seen:
            continue
        seen.add(name)
        res.append("kernel: " + name[:100])
        for key, name in WANT:
            if key in hdr:
                i = hdr.index(key)
                res.append("  %-20s %18s %s" % (name, r[i], units[i]))
    return "\n".join(res)


if __name__ == "__main__":
    for rep in sys.argv[1:]:
        print("==", rep)
        print(summarize(rep))
