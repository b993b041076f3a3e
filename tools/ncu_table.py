#!/usr/bin/env python
"""Compact per-kernel table from an `ncu --page raw --csv` export:  python tools/ncu_table.py file.csv [name-filter]"""
import csv
import sys

SHORT = {
    "t_ms": "gpu__time_duration.sum", "regs": "launch__registers_per_thread", "grid": "launch__grid_size", "blk": "launch__block_size",
    "occ%": "sm__warps_active.avg.pct_of_peak_sustained_active", "inst_G": "smsp__inst_executed.sum",
    "alu%": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "fma%": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "lsu%": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "issue%": "smsp__issue_active.avg.pct_of_peak_sustained_active", "rdGB": "dram__bytes_read.sum", "wrGB": "dram__bytes_write.sum",
    "icc%": "sm__icc_request_hit_rate.pct",
    "noinst": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "longsb": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "shortsb": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "mathpipe": "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "dispatch": "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "mio": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "lg": "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "notsel": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    print("kernel".ljust(44), " ".join(k.rjust(8) for k in SHORT))
    for r in data:
        if flt not in r[ki]:
            continue
        vals = []
        for k, n in SHORT.items():
            try:
                v = float(r[hdr.index(n)].replace(",", ""))
                u = units[hdr.index(n)]
                if k == "inst_G":
                    v /= 1e9
                if k in ("rdGB", "wrGB"):
                    v *= {"Gbyte": 1, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "Tbyte": 1e3}.get(u, 1)
                if k == "t_ms":
                    v *= {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1)
                vals.append(f"{v:8.2f}")
            except (ValueError, IndexError):
                vals.append("       -")
        name = r[ki].replace("void nb::", "").replace("nb::", "")
        print(name[:44].ljust(44), " ".join(vals))


if __name__ == "__main__":
    main()
