#!/usr/bin/env python
"""Copy the outputs of tools/profile_round.sh <tag> (gpurun_out/) into profiles/ and regenerate the summaries:
<tag>_bench.json (+ _reference, _cfg5, _cli*), <tag>_launches.{csv,md}, <tag>_ncu_summary.md, <tag>_ncu_hbm_kernels.md (+ _cfg5),
<tag>_ncu_fused_layers.md, <tag>_batch_sweep.md, <tag>_sass_summary.md and the per-launch DRAM traffic of the kernels in
traffic_per_launch.json.      usage: publish_profiles.py [tag]   (default r02)"""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
G, P = "gpurun_out", "profiles"
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
for name in ("bench.json", "bench_reference.json", "bench_cfg5.json", "launches.csv", "batch_sweep.md", "cli.json", "cli_fastq.json", "cli_cfg1.json"):
    if os.path.exists("%s/%s_%s" % (G, TAG, name)):
        shutil.copy("%s/%s_%s" % (G, TAG, name), "%s/%s_%s" % (P, TAG, name))
run = lambda *a, **k: subprocess.run(list(a), capture_output=True, text=True, **k).stdout
run(sys.executable, "profiles/summarize_ncu.py", G + "/%s_all.ncu-rep" % TAG, P + "/%s_ncu_summary.md" % TAG)
run(sys.executable, "profiles/summarize_ncu.py", G + "/%s_hbm.ncu-rep" % TAG, P + "/%s_ncu_hbm_kernels.md" % TAG)
if os.path.exists(G + "/%s_cnn.ncu-rep" % TAG):
    run(sys.executable, "profiles/summarize_ncu.py", G + "/%s_cnn.ncu-rep" % TAG, P + "/%s_ncu_cnn.md" % TAG)
if os.path.exists(G + "/%s_hbm_cfg5.ncu-rep" % TAG):
    run(sys.executable, "profiles/summarize_ncu.py", G + "/%s_hbm_cfg5.ncu-rep" % TAG, P + "/%s_ncu_hbm_kernels_cfg5.md" % TAG)
run(sys.executable, "tools/sass_summary.py", P + "/%s_sass_summary.md" % TAG)
d = json.loads(open(P + "/%s_bench.json" % TAG).read().strip().splitlines()[-1])
st, tot, r = d["stage_ms_per_step"], d["ms_per_step"], d["roofline"]
side = {"lstm0": "read_rnn1 of the next chunk, under rec2", "heads": "heads tail, under the next chunk's kernels"}
L = ["# ncu launch list summary (%s)" % TAG, "",
     "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv python bench.py --steps 2 --warmup 1 --reads-per-step 16 "
     "--no-cpu-baseline` (raw: %s_launches.csv; cold-cache, serialised: compare SHARES)" % TAG, "",
     run(sys.executable, "tools/launch_summary.py", G + "/%s_launches.csv" % TAG).strip(), "",
     "## CUDA-event stage times of the same build (`python bench.py --steps 5 --warmup 3`, 128 reads/step; profiles/%s_bench.json)" % TAG, "",
     "`lstm0` and `heads` run on a low-priority side stream (see DESIGN.md section 4): their event times are stretched durations that overlap other "
     "stages; shares are of ms_per_step = %.1f ms." % tot, "", "| stage | ms/step | share of step |", "|---|---|---|"]
for k, v in sorted(st.items(), key=lambda kv: -kv[1]):
    L.append("| %s | %.2f | %.1f %%%s |" % (k, v, 100 * v / tot, " (overlapped: %s)" % side[k] if k in side else ""))
L += ["", "value = %.4g bases/s, e2e = %.4g bases/s; dominant stage %s = %s: %.1f TFLOP/s algorithmic = %.1f %% of bf16 sustained (%.1f %% of the 3-pass "
      "ceiling); whole path %.1f TFLOP/s = %.1f %%; clocks %s" % (d["value"], d["e2e"]["value"], r["stage"], r["kernel"].split(" (")[0], r["achieved"],
      100 * r["frac"], 300 * r["frac"], d["whole_path"]["achieved_tflops"], 100 * d["whole_path"]["frac_of_bf16_sustained"], json.dumps(d["clocks"])),
      "", "Stage `proj2` is only the CNN-feature gather of the window tiles with a read boundary (`tile_base_kernel` + `gather_sig_kernel`); `rec2` / `rec3` are the fused layer kernels (projection + recurrence).",
      "Shares: CUDA events total_rnn1 / total_rnn2 / read_rnn11 = %.1f %% / %.1f %% / %.1f %% of the step; compare the serialised ncu list above."
      % (100 * st["rec2"] / tot, 100 * st["rec3"] / tot, 100 * st["rec1"] / tot)]
open(P + "/%s_launches.md" % TAG, "w").write("\n".join(L) + "\n")
# per-launch DRAM traffic from the --set full capture
raw = run("ncu", "-i", G + "/%s_all.ncu-rep" % TAG, "--page", "raw", "--csv")
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
tb = lambda v, u: float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
acc = {}
for x in data:
    name = x[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("nrv::", "").replace(" ", "")
    import re as _re
    name = _re.sub(r"^(lstm_fused_pair_kernel<\d+,\d+)(,\d)+>", r"\1>", name).replace("<unnamed>::", "")       # PF8 / RF8 / OUT8 template flags
    t = tb(x[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + tb(x[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    acc.setdefault(name, []).append(t)
mean = {k: sum(v) / len(v) for k, v in acc.items()}
stage_of = {"rec2": "lstm_fused_pair_kernel<192,128>", "rec3": "lstm_fused_pair_kernel<256,64>",
            "rec1": "lstm_fused_tc64_pp_kernel" if "lstm_fused_tc64_pp_kernel" in mean else "lstm_fused_tc64_kernel",
            "lstm0": "read_rnn1_kernel", "heads_gemm": "gemm_f16x3_kernel<128,1>", "heads": "heads_tail_thread_kernel<11>"}
t = json.load(open(P + "/traffic_per_launch.json"))
for s_, k in stage_of.items():
    if k in mean:
        t[s_] = {"kernel": k, "dram_bytes_per_launch": round(mean[k])}
if "gather_sig_kernel" in mean:
    t["proj2"] = {"kernel": "gather_sig_kernel", "dram_bytes_per_launch": round(mean["gather_sig_kernel"])}
t["_source"] = ("ncu --set full --clock-control none, profiles/%s_ncu_summary.md (gpurun_out/%s_all.ncu-rep); one launch = one chunk of 151,552 windows "
                "of one model (the split-path figures under _split_path are per 37,888-window chunk, round 1). total_rnn2 reads its input once per "
                "DIRECTION (fwd and bwd clusters walk t in opposite order); total_rnn1 takes the CNN features from the per-base table (L2 hits)" % (TAG, TAG))
json.dump(t, open(P + "/traffic_per_launch.json", "w"), indent=1)
# fused-layer stall summaries
out = ["# per-source-line stall samples of the fused layer kernels (ncu --set full, gpurun_out/%s_all.ncu-rep; "
       "tools/ncu_top.py + tools/ncu_lines.py)" % TAG]
def rep(want):                     # cnn_kernel runs before the model kernels of a step: its own capture (tools/prof_cnn.sh)
    return G + ("/%s_cnn.ncu-rep" if "cnn" in want else "/%s_all.ncu-rep") % TAG
for title, want, sub in (("<192,128,OUT8> (total_rnn1)", "(int)192", "ILi192ELi128ELb0ELb1E"), ("<256,64,F8> (total_rnn2)", "(int)256", "ILi256ELi64ELb1ELb0E"),
                         ("cnn_kernel (K2)", "cnn_kernel", "cnn_kernel")):
    env = dict(os.environ, NCU_KERNEL=want, NCU_FILTER="--kernel-name-base demangled --kernel-name regex:" + ("cnn_kernel" if "cnn" in want else want.replace("(int)", "")))
    out += ["", "## " + title, "```",
            "\n".join(run(sys.executable, "tools/ncu_top.py", rep(want), "0", env=env).split("\n")[:22]),
            "\n".join(l[:200] for l in run(sys.executable, "tools/ncu_lines.py", rep(want), sub, "22", env=env).split("\n")), "```"]
open(P + "/%s_ncu_fused_layers.md" % TAG, "w").write("\n".join(out) + "\n")
print("value %.4g e2e %.4g" % (d["value"], d["e2e"]["value"]), st)
print({k: v for k, v in t.items() if not k.startswith("_")})
