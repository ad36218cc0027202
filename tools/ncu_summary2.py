#!/usr/bin/env python
"""Prints the handful of ncu metrics the tick-kernel optimisation log uses from a .ncu-rep (run where ncu is installed):
python tools/ncu_summary2.py report.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("==", d.get("Kernel Name", "?")[:100])
        for k in WANT:
            if k in d:
                print(f"  {k} = {d[k]}")
        st = sorted(((float(v), k[len(STALL):-len('_per_issue_active.ratio')]) for k, v in d.items()
                     if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)
        print("  stalls per issue:", ", ".join(f"{k} {v:.2f}" for v, k in st[:8]))


if __name__ == "__main__":
    main()
