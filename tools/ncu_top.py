#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel): headline metrics, stall reasons, top stalled SASS lines.  usage: ncu_top.py rep [n]"""
import os, collections, csv, io, re, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + os.environ.get("NCU_FILTER", "").split(), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
          "lts__t_bytes.sum.per_second", "l1tex__data_bank_conflicts_pipe_lsu.sum"]:
    if k in hdr:
        print("%-75s %s %s" % (k, vals[hdr.index(k)], rows[1][hdr.index(k)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + os.environ.get("NCU_FILTER", "").split(), capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp] or 0) for r in data)
print("total samples", tot)
for name in hdr:
    if name.startswith("stall_") and "Not" not in name:
        v = sum(int(r[hdr.index(name)] or 0) for r in data)
        if v * 50 > tot:
            print("  %-24s %6d  %.1f %%" % (name, v, 100.0 * v / tot))
for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:n]:
    print(r[isamp].rjust(6), r[iex].rjust(9), r[isrc][:110])
