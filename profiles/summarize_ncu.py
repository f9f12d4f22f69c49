"""Summarise an ncu report (--set full) into a markdown table + traffic_per_launch.json.
usage: python profiles/summarize_ncu.py gpurun_out/prof_r01_all.ncu-rep profiles/r01_ncu_summary.md"""
import csv
import io
import json
import subprocess
import sys

rep, out_md = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "XU %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


lines = ["| kernel | " + " | ".join(n for _, n in want) + " | DRAM GB/s |", "|---|" + "---|" * (len(want) + 1)]
traffic = {}
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("nrv::", "")
    vals = []
    for k, _ in want:
        v, u = r[col[k]], units[col[k]]
        if k == "gpu__time_duration.sum":      # ncu picks the unit per report: always print microseconds
            vals.append("%.1f" % (float(v.replace(",", "")) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3,
                                                                 "s": 1e6, "second": 1e6}.get(u, 1)))
        elif k.startswith("dram__bytes"):
            vals.append("%.3f GB" % (to_bytes(v, u) / 1e9))
        else:
            try:
                vals.append("%.1f" % float(v.replace(",", "")))
            except ValueError:
                vals.append(v)
    t_us = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
    t_us *= {"ns": 1e-3, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(units[col["gpu__time_duration.sum"]], 1)
    tot = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
        to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    lines.append("| %s | %s | %.0f |" % (name, " | ".join(vals), tot / t_us / 1e3))
    traffic.setdefault(name + " grid=" + r[col["launch__grid_size"]], []).append(tot)
open(out_md, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(out_md.replace(".md", "_traffic.json"), "w"), indent=1)
