#!/usr/bin/env python
"""Reduce an .ncu-rep (ncu --set full) to the handful of metrics DESIGN.md / bench.py quote, one line per metric.
usage: python tools/ncu_summary.py file.ncu-rep [header text ...]  > profiles/xxx.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main():
    rep = sys.argv[1]
    for h in sys.argv[2:]:
        print("# " + h)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[hdr], rows[hdr + 1]
    for data in rows[hdr + 2:]:
        if len(data) != len(names):
            continue
        d = dict(zip(names, data))
        u = dict(zip(names, units))
        print(f"{'Kernel Name':90s} {d['Kernel Name']}")
        for m in WANT:
            if m in d:
                print(f"{m:90s} {d[m]} {u.get(m, '')}")
        try:
            rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
            tot = rd * scale.get(u["dram__bytes_read.sum"], 1) + wr * scale.get(u["dram__bytes_write.sum"], 1)
            print(f"{'dram bytes read + written (per launch)':90s} {tot:.0f} byte")
        except Exception:
            pass
        print()


if __name__ == "__main__":
    main()
