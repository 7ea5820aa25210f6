#!/usr/bin/env python
"""Summarise an .ncu-rep (run where ncu is installed): python tools/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum", "sm__cycles_elapsed.avg.per_second",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_active.avg", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:80], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
    for k in KEYS:
        hits = [i for i, h in enumerate(hdr) if h == k or h.endswith("." + k)]
        for i in hits[:1]:
            print(f"   {k:70s} {r[i]:>18s} {units[i]}")
    tp = [i for i, h in enumerate(hdr) if "pipe_tensor" in h]
    for i in tp:
        if r[i] not in ("", "0", "n/a"):
            print(f"   {hdr[i]:70s} {r[i]:>18s} {units[i]}")
