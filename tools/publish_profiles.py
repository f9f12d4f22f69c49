#!/usr/bin/env python
"""Copy the outputs of tools/profile_round.sh (gpurun_out/) into profiles/ and regenerate the summaries:
r01_bench_final.json, r01_launches_final.{csv,md}, r01_ncu_summary.md, r01_ncu_hbm_kernels.md, r01_ncu_fused_layers.md,
r01_batch_sweep.md and the per-launch DRAM traffic of the kernels in traffic_per_launch.json."""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
G, P = "gpurun_out", "profiles"
shutil.copy(G + "/r01_bench.json", P + "/r01_bench_final.json")
shutil.copy(G + "/r01_launches.csv", P + "/r01_launches_final.csv")
shutil.copy(G + "/r01_batch_sweep.md", P + "/r01_batch_sweep.md")
run = lambda *a, **k: subprocess.run(list(a), capture_output=True, text=True, **k).stdout
run(sys.executable, "profiles/summarize_ncu.py", G + "/r01_all.ncu-rep", P + "/r01_ncu_summary.md")
run(sys.executable, "profiles/summarize_ncu.py", G + "/r01_hbm.ncu-rep", P + "/r01_ncu_hbm_kernels.md")
d = json.loads(open(P + "/r01_bench_final.json").read().strip().splitlines()[-1])
st, tot, r = d["stage_ms_per_step"], d["ms_per_step"], d["roofline"]
side = {"lstm0": "read_rnn1 of the next chunk, under rec2", "heads": "heads tail, under the next chunk's kernels"}
L = ["# ncu launch list summary (round 1, final build)", "",
     "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv python bench.py --steps 2 --warmup 1 --reads-per-step 16 "
     "--no-cpu-baseline` (raw: r01_launches_final.csv; cold-cache, serialised: compare SHARES)", "",
     run(sys.executable, "tools/launch_summary.py", G + "/r01_launches.csv").strip(), "",
     "## CUDA-event stage times of the same build (`python bench.py --steps 5 --warmup 3`, 128 reads/step; profiles/r01_bench_final.json)", "",
     "`lstm0` and `heads` run on a low-priority side stream (see DESIGN.md section 4): their event times are stretched durations that overlap other "
     "stages; shares are of ms_per_step = %.1f ms." % tot, "", "| stage | ms/step | share of step |", "|---|---|---|"]
for k, v in sorted(st.items(), key=lambda kv: -kv[1]):
    L.append("| %s | %.2f | %.1f %%%s |" % (k, v, 100 * v / tot, " (overlapped: %s)" % side[k] if k in side else ""))
L += ["", "value = %.4g bases/s, e2e = %.4g bases/s; dominant stage %s = %s: %.1f TFLOP/s algorithmic = %.1f %% of bf16 sustained (%.1f %% of the 3-pass "
      "ceiling); whole path %.1f TFLOP/s = %.1f %%; clocks %s" % (d["value"], d["e2e"]["value"], r["stage"], r["kernel"].split(" (")[0], r["achieved"],
      100 * r["frac"], 300 * r["frac"], d["whole_path"]["achieved_tflops"], 100 * d["whole_path"]["frac_of_bf16_sustained"], json.dumps(d["clocks"])),
      "", "Stage `proj2` is only the CNN-feature gather (`gather_sig_kernel`); `rec2` / `rec3` are the fused layer kernels (projection + recurrence).",
      "Shares: CUDA events total_rnn1 / total_rnn2 / read_rnn11 = %.1f %% / %.1f %% / %.1f %% of the step; compare the serialised ncu list above."
      % (100 * st["rec2"] / tot, 100 * st["rec3"] / tot, 100 * st["rec1"] / tot)]
open(P + "/r01_launches_final.md", "w").write("\n".join(L) + "\n")
# per-launch DRAM traffic from the --set full capture
raw = run("ncu", "-i", G + "/r01_all.ncu-rep", "--page", "raw", "--csv")
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
tb = lambda v, u: float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
acc = {}
for x in data:
    name = x[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("nrv::", "").replace(" ", "")
    t = tb(x[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + tb(x[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    acc.setdefault(name, []).append(t)
mean = {k: sum(v) / len(v) for k, v in acc.items()}
stage_of = {"rec2": "lstm_fused_pair_kernel<192,128>", "rec3": "lstm_fused_pair_kernel<256,64>", "rec1": "lstm_fused_tc64_kernel",
            "lstm0": "read_rnn1_kernel", "heads_gemm": "gemm_f16x3_kernel<128,1>", "heads": "heads_tail_thread_kernel<11>"}
t = json.load(open(P + "/traffic_per_launch.json"))
for s_, k in stage_of.items():
    if k in mean:
        t[s_] = {"kernel": k, "dram_bytes_per_launch": round(mean[k])}
json.dump(t, open(P + "/traffic_per_launch.json", "w"), indent=1)
# fused-layer stall summaries
out = ["# per-source-line stall samples of the fused layer kernels (ncu --set full of the final round-1 build, gpurun_out/r01_all.ncu-rep; "
       "tools/ncu_top.py + tools/ncu_lines.py with NCU_FILTER)"]
for i, (title, sub) in enumerate((("<192,128> (total_rnn1)", "ILi192ELi128E"), ("<256,64> (total_rnn2)", "ILi256ELi64E"))):
    env = dict(os.environ, NCU_FILTER="--kernel-name regex:lstm_fused_pair_kernel --launch-skip %d --launch-count 1" % i)
    out += ["", "## lstm_fused_pair_kernel" + title, "```",
            "\n".join(run(sys.executable, "tools/ncu_top.py", G + "/r01_all.ncu-rep", "0", env=env).split("\n")[:22]),
            "\n".join(l[:200] for l in run(sys.executable, "tools/ncu_lines.py", G + "/r01_all.ncu-rep", sub, "22", env=env).split("\n")), "```"]
open(P + "/r01_ncu_fused_layers.md", "w").write("\n".join(out) + "\n")
print("value %.4g e2e %.4g" % (d["value"], d["e2e"]["value"]), st)
print({k: v for k, v in t.items() if not k.startswith("_")})
