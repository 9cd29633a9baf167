#!/usr/bin/env python
"""Summarise one .ncu-rep (raw page) into the handful of metrics the roofline discussion needs."""
import csv, subprocess, sys, json
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg"]
for vals in rows[2:]:
    d = {}
    for i, h in enumerate(hdr):
        if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            try:
                v = float(vals[i].replace(",", ""))
            except ValueError:
                v = vals[i]
            if isinstance(v, float) and "issue_stalled" in h and v < 0.3:
                continue
            d[h] = (v, units[i])
    for k, (v, u) in d.items():
        print(f"{k:85s} {u:14s} {v}")
    print("-" * 60)
