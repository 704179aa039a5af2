#!/usr/bin/env python
"""Summarise .ncu-rep captures into small text files for profiles/ (the reports themselves are
multi-MB and stay in gpurun_out/).  Usage: summarize_ncu.py out.txt rep1.ncu-rep [rep2 ...]"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    lines = []
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            lines.append(f"== {rep}: unreadable\n")
            continue
        hdr, units = rows[0], rows[1]
        for d in rows[2:]:
            name = d[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            lines.append(f"== {rep} :: {name[:150]}")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    lines.append(f"   {w:72s} {d[i]:>18s} {units[i]}")
            for i, h in enumerate(hdr):
                if "issue_stalled" in h and "per_issue_active" in h:
                    try:
                        v = float(d[i])
                    except ValueError:
                        continue
                    if v >= 0.08:
                        lines.append(f"   stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):66s} {v:18.3f}")
            lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out, len(lines), "lines")


if __name__ == "__main__":
    main()
