#!/usr/bin/env python
"""Text summary of an ncu report for profiles/: selected raw metrics + stall samples per reason / opcode / hottest
instructions (scripts/ncu_hotspots.py).   python scripts/ncu_summary.py report.ncu-rep "header line" > profiles/x.txt"""
import csv
import io
import os
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "local_load", "local_store")


def main():
    rep, header = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    here = os.path.dirname(os.path.abspath(__file__))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    print(header)
    print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for i, name in enumerate(hdr):
        if name in KEEP or name.startswith("smsp__average_warps_issue_stalled") and name.endswith("_per_issue_active.ratio") or "local_op" in name and name.endswith(".sum"):
            print(f"{name} [{units[i]}] = {vals[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], check=True, capture_output=True, text=True).stdout
    tmp = rep + ".source.csv"
    open(tmp, "w").write(src)
    print("\n== stall samples (scripts/ncu_hotspots.py on --page source) ==")
    print(subprocess.run([sys.executable, os.path.join(here, "ncu_hotspots.py"), tmp, "24"], check=True, capture_output=True, text=True).stdout)
    os.remove(tmp)


if __name__ == "__main__":
    main()
