#!/usr/bin/env python
"""bench.py -- revised bases/s of the revision-inference hot path on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads-per-step R]

A *step* is one pass of the whole hot path (K1 segmentation -> K2 CNN -> K3 Bi-LSTM x2 models + heads
-> K4 decode) over one ragged slab of synthetic reads of BASELINE.json's configs[1] shape ("ecoli
model, synthetic 100k reads x 10 kb (4 kHz signal), 1xB200"): R reads x 10,000 bases per step, i.e.
the steady state of the 100k-read job (a full 1e9-base pass would not fit the bench budget).

  value : whole-job revised bases/s, inputs already resident in HBM (nrv_revise_batch_device),
          timed with CUDA events on the library's stream, max over ranks.
  e2e   : same metric through the reference-facing call with HOST buffers (nrv_revise_batch: pinned
          host -> device copies, kernels, device -> host copy of the revised sequences), per step.
  roofline     : dominant kernel = Bi-LSTM layer `total_rnn1` (57.7 % of the algorithmic FLOPs);
                 achieved = algorithmic FLOPs per launch / CUDA-event duration of those launches.
  cpu_baseline : the CPU oracle (numpy fp32 restatement, all host cores) on a bounded sample.

--impl reference times the reference's own CPU algorithm for the path (the oracle port with the
reference-faithful per-read python segmentation and per-window CNN recompute; Keras/TF cannot be
installed here) on the host cores, same metric/unit/config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv[1:]:
    # the reference arm uses every host core; torchrun exports OMP_NUM_THREADS=1, which would pin the BLAS to one thread
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "revised_bases_per_sec"
UNIT = "bases/s"
READ_LEN = 10_000
# algorithmic MACs per window, per model; CNN is per base
MAC_LSTM = (30_976, 540_672, 3_604_480, 1_802_240)
MAC_HEADS = 228_448 + 96          # model1; model2 has 80 in the last layer
MAC_CNN = 36_400
FLOP_PER_BASE = 24.97e6


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle_bases_per_sec(n_reads: int, read_len: int, faithful: bool, seed: int = 99):
    """Time the CPU oracle on a bounded sample of the same workload.  Returns (bases/s, seconds, bases)."""
    from nanoreviser_b200 import synth, weights
    from oracle import nanorev_oracle as orc
    m1, m2 = weights.load_species("ecoli", os.path.join(ROOT, "model"))
    b = synth.make_batch([read_len] * n_reads, seed=seed)
    t0 = time.perf_counter()
    total = 0
    for i in range(b.n_reads):
        s0, s1 = int(b.sig_off[i]), int(b.sig_off[i + 1])
        b0, b1 = int(b.base_off[i]), int(b.base_off[i + 1])
        starts = b.starts[b0:b1].astype(np.int64)
        length = np.diff(np.append(starts, starts[-1] + b.last_dur[i])).astype(float)
        bases = [chr(c) for c in b.bases[b0:b1]]
        if faithful:
            # what signal_segmentation + model.predict([S, X]) do: per-base python loop, 11x window
            # materialisation, CNN recomputed for every timestep of every window
            win, mean, std, shift, scale = orc.signal_segmentation(b.signal[s0:s1], starts, int(b.last_dur[i]))
            x = orc.feature_columns(bases, mean, std, shift, scale, length, b.ev_mean[b0:b1], b.ev_std[b0:b1])
            X, S = orc.make_windows(x, win, m1.window)
            P1 = np.concatenate([orc.forward_windows(m1, S[a:a + 2048], X[a:a + 2048]) for a in range(0, len(X), 2048)])
            P2 = np.concatenate([orc.forward_windows(m2, S[a:a + 2048], X[a:a + 2048]) for a in range(0, len(X), 2048)])
            M = len(X)
            orc.get_base_1(bases[5:5 + M], P1.argmax(1), P2.argmax(1) + 2)
        else:
            res = orc.revise_arrays(m1, m2, bases, starts, length, b.signal[s0:s1], b.ev_mean[b0:b1], b.ev_std[b0:b1])
            if i == 0:
                cpu_oracle_bases_per_sec.last = (b, res["revised"])     # checker for the GPU result on the same read
        total += b1 - b0
    dt = time.perf_counter() - t0
    return total / dt, dt, total


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n = max(1, args.ref_reads)
    vals = []
    for _ in range(max(1, args.warmup > 0)):
        cpu_oracle_bases_per_sec(1, 1000, True)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, dt, nb = cpu_oracle_bases_per_sec(n, args.ref_read_len, True)
        vals.append((v, dt, nb))
    v = sum(x[2] for x in vals) / sum(x[1] for x in vals)
    sample = "%d synthetic cfg2-shape read(s) x %d bases per step, %d steps" % (n, args.ref_read_len, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(x[1] for x in vals) / len(vals),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg2: ecoli model, synthetic 10 kb reads (4 kHz signal); bounded CPU sample of the slab",
                       "reads_per_step": n, "bases_per_step": n * args.ref_read_len},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "oracle port of the reference algorithm (numpy fp32 + OpenBLAS, per-read python "
                                     "segmentation, per-window CNN recompute); Keras 2.2.4/TF 1.12 not installable"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads-per-step", type=int, default=128)
    ap.add_argument("--species", default="ecoli")
    ap.add_argument("--ref-reads", type=int, default=1)
    ap.add_argument("--ref-read-len", type=int, default=4000)
    ap.add_argument("--cpu-sample-bases", type=int, default=10_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from nanoreviser_b200 import engine, synth, weights

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peaks, peak_src = load_peaks()

    m1, m2 = weights.load_species(args.species, os.path.join(ROOT, "model"))
    rv = engine.Reviser(m1, m2, device=local_rank)
    W = rv.window
    R = args.reads_per_step
    # two distinct slabs per rank, alternated, so that no step re-reads the previous step's inputs from L2
    slabs = [synth.make_batch([READ_LEN] * R, seed=1000 + 17 * rank + s) for s in range(2)]
    n_bases = slabs[0].n_bases
    n_win = slabs[0].n_windows(W)
    cap = 2 * n_bases + R + 16
    stream = torch.cuda.ExternalStream(rv.stream, device=local_rank)

    # ---- device-resident copies (torch only as the allocator / stream plumbing) ---------------------
    dev = torch.device("cuda", local_rank)
    dslabs = []
    for b in slabs:
        t = {k: torch.from_numpy(getattr(b, k)).to(dev) for k in ("signal", "starts", "bases", "ev_mean", "ev_std", "last_dur")}
        dslabs.append(t)
    d_rev = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_off = torch.empty(R + 1, dtype=torch.int64, device=dev)
    d_status = torch.empty(R, dtype=torch.int32, device=dev)
    dres = {"revised": d_rev.data_ptr(), "out_off": d_off.data_ptr(), "status": d_status.data_ptr()}
    torch.cuda.synchronize()

    def step_device(i):
        b, t = slabs[i & 1], dslabs[i & 1]
        rv.revise_batch_device(R, b.sig_off, b.base_off, {k: v.data_ptr() for k, v in t.items()}, dres, cap)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ------------------------------------------------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    rv.synchronize()
    # ---- timed region: `value` ---------------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    rv.set_stage_timing(True)
    launches0 = rv.launch_count
    barrier()
    with torch.cuda.stream(stream):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            step_device(i)
        e1.record(stream)
    rv.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    gpu_launches = rv.launch_count - launches0
    stage_ms = rv.stage_ms()
    stage_launches = rv.stage_launches()
    rv.set_stage_timing(False)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * n_bases * args.steps / (ms_max * 1e-3)

    # ---- e2e: host buffers through nrv_revise_batch (H2D + kernels + D2H inside the timed region) -----
    pinned = []
    for b in slabs:
        pb = engine.Batch(**{k: torch.from_numpy(getattr(b, k)).pin_memory().numpy() for k in
                             ("signal", "sig_off", "starts", "base_off", "bases", "ev_mean", "ev_std", "last_dur")})
        pinned.append(pb)
    out = engine.ReviseResult(torch.empty(cap, dtype=torch.uint8).pin_memory().numpy(),
                              torch.empty(R + 1, dtype=torch.int64).pin_memory().numpy(),
                              torch.empty(R, dtype=torch.int32).pin_memory().numpy())
    for i in range(max(1, min(args.warmup, 2))):
        rv.revise_batch(pinned[i & 1], out=out)
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for i in range(args.steps):
        rv.revise_batch(pinned[i & 1], out=out)
        d2h += int(out.out_off[-1]) + out.out_off.nbytes + out.status.nbytes
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * n_bases * args.steps / float(t_e.item())
    # sanity: the device-resident and the host path produce the same bytes
    chk = rv.revise_batch(pinned[(args.steps - 1) & 1])
    torch.cuda.synchronize()
    same = bool(np.array_equal(d_off.cpu().numpy(), chk.out_off)) and \
        bool(np.array_equal(d_rev.cpu().numpy()[:chk.out_off[-1]], chk.revised[:chk.out_off[-1]]))

    if rank == 0:
        # ---- roofline: per model kernel, algorithmic FLOPs / CUDA-event time ----------
        peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
        tc_path = stage_ms.get("proj2", 0.0) > 0.0
        # algorithmic MACs per window and per model of each stage (both directions, 11 timesteps)
        # which kernels ran for total_rnn1 / total_rnn2 (library defaults: both fused; NRV_TRNN1 / NRV_TRNN2 = split select GEMM + recurrence)
        fused1 = tc_path and os.environ.get("NRV_TRNN1", "fused") != "split"
        fused2 = tc_path and stage_ms.get("proj3", 0.0) == 0.0
        if tc_path:
            macs = {"lstm0": 30_976, "proj1": 180_224, "rec1": 360_448, "proj2": 2_162_688, "rec2": 1_441_792,
                    "proj3": 1_441_792, "rec3": 360_448, "heads_gemm": 225_280, "heads": 3_264}
            names = {"proj2": "gemm_f16x3_pair_kernel (total_rnn1 input projection, tcgen05 cta_group::2, 3-pass split-fp16)",
                     "rec2": "lstm_rec_tc128_pair_kernel (total_rnn1 recurrence, tcgen05 cta_group::2)",
                     "proj3": "gemm_f16x3_pair_kernel (total_rnn2 input projection, tcgen05 cta_group::2)",
                     "rec3": "lstm_rec_tc64_kernel (total_rnn2 recurrence, tcgen05)",
                     "proj1": "gemm_f16x3_kernel<256> (read_rnn11 input projection, tcgen05)",
                     "rec1": "lstm_fused_tc64_kernel (read_rnn11 projection+recurrence, tcgen05)",
                     "heads_gemm": "gemm_f16x3_kernel<128,FUSE2> (dense head 128->128->32, tcgen05)",
                     "heads": "heads_tail_thread_kernel (dense 32->6, flatten, feature, softmax, argmax; fp32 SIMT)",
                     "lstm0": "read_rnn1_kernel (read_rnn1, fp32 SIMT, one thread per window pair)"}
            if fused1:      # stage proj2 holds only the CNN-feature gather (no MACs); rec2 is the whole layer
                macs.pop("proj2")
                macs["rec2"] = 3_604_480
                names["rec2"] = ("lstm_fused_pair_kernel<192,128> (total_rnn1 projection+recurrence fused, tcgen05 cta_group::2, "
                                 "cluster of 4, h in TMEM, 3-pass split-fp16)")
            if fused2:
                macs.pop("proj3")
                macs["rec3"] = 1_802_240
                names["rec3"] = ("lstm_fused_pair_kernel<256,64> (total_rnn2 projection+recurrence fused, tcgen05 cta_group::2, "
                                 "h in TMEM, 3-pass split-fp16)")
        else:
            macs = {"lstm0": 30_976, "rec1": 540_672, "rec2": 3_604_480, "rec3": 1_802_240, "heads": 228_544}
            names = {k: "lstm_layer_kernel (fp32 SIMT, fused projection+recurrence)" for k in macs}
            names["heads"] = "heads_kernel (fp32 SIMT)"
        # algorithmic HBM bytes per window and per model of each stage (activations
        # are fp16 hi/lo pairs = 4 B per value, pre-activations fp32; 11 timesteps; weights are L2-resident)
        abytes = {"lstm0": 11 * (6 + 32) * 4, "rec1": 11 * (32 + 128) * 4, "proj2": 11 * (192 + 1024) * 4,
                  "rec2": 11 * (1024 + 256) * 4, "proj3": 11 * (256 + 512) * 4, "rec3": 11 * (512 + 128) * 4,
                  "heads_gemm": 11 * (128 + 32) * 4, "heads": 11 * 32 * 4 + 32} if tc_path else {}
        if fused1:          # reads x = [read_rnn11 | CNN features] (192), writes h (256): no pre-activations
            abytes["rec2"] = 11 * (192 + 256) * 4
        if fused2:
            abytes["rec3"] = 11 * (256 + 128) * 4
        hbm_peak = float(peaks.get("hbm_gbs"))
        traffic_db = {}
        tp = os.path.join(ROOT, "profiles", "traffic_per_launch.json")
        if os.path.exists(tp):
            traffic_db = json.load(open(tp))
        kernels = {}
        for k, mac in macs.items():
            t_ms = stage_ms.get(k, 0.0)
            if t_ms <= 0:
                continue
            fl = 2.0 * mac * n_win * 2 * args.steps          # x2 models
            ach = fl / (t_ms * 1e-3) / 1e12
            nl = max(stage_launches.get(k, 1), 1)
            kernels[k] = {"kernel": names.get(k, k), "achieved": ach, "frac": ach / peak, "launches": int(nl),
                          "avg_launch_ms": t_ms / nl, "flops_per_launch": fl / nl, "share_of_step": t_ms / ms if ms > 0 else None}
            if k in abytes:
                gbs = abytes[k] * n_win * 2 * args.steps / (t_ms * 1e-3) / 1e9
                kernels[k].update(hbm_bytes_per_launch=abytes[k] * n_win * 2 * args.steps / nl, hbm_achieved_gbs=gbs,
                                  hbm_frac=gbs / hbm_peak)
        # read_rnn1 (lstm0) of the next (chunk, model) runs on a low-priority side stream UNDER the fused total_rnn1 kernel (on the
        # SMs its clusters of 4 cannot use): its event time is its stretched duration there, not exclusive time
        overlapped = fused1 and os.environ.get("NRV_OVERLAP", "1") != "0"
        side = ("lstm0", "heads") if overlapped else ()          # read_rnn1 under rec2; heads tail under the next iteration's kernels
        for k in side:
            if k in kernels:
                kernels[k]["overlapped_with"] = "rec2" if k == "lstm0" else "rec1 (next chunk)"
        dom = max((k for k in kernels if k not in side), key=lambda k: stage_ms[k])
        roofline = {"bound": "tensor", "kernel": kernels[dom]["kernel"], "stage": dom, "achieved": kernels[dom]["achieved"],
                    "peak": peak, "unit": "TFLOP/s", "frac": kernels[dom]["frac"],
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s)" % peak_src,
                    "launches": kernels[dom]["launches"], "avg_launch_ms": kernels[dom]["avg_launch_ms"],
                    "flops_per_launch": kernels[dom]["flops_per_launch"],
                    # measured DRAM bytes per launch (ncu --set full) -- only if the capture is of the kernel that ran
                    "traffic": (traffic_db.get(dom, {}).get("dram_bytes_per_launch")
                                if kernels[dom]["kernel"].startswith(str(traffic_db.get(dom, {}).get("kernel", "?")).split()[0].split("<")[0])
                                else None),
                    "share_of_step": kernels[dom]["share_of_step"],
                    "hbm": {"achieved": kernels[dom].get("hbm_achieved_gbs"), "peak": hbm_peak, "unit": "GB/s",
                            "frac": kernels[dom].get("hbm_frac"), "bytes_per_launch": kernels[dom].get("hbm_bytes_per_launch")},
                    "note": ("algorithmic FLOPs (1 pass); the kernel spends 3 fp16 MMA passes per product to stay fp32-equivalent "
                             "(effective tensor ceiling = peak / 3) and runs under the 1000 W power cap; fused layers move only "
                             "x and h through HBM (see traffic)") if (fused1 or fused2) else
                            ("algorithmic FLOPs (1 pass); the kernel spends 3 fp16 MMA passes per product to stay "
                             "fp32-equivalent, and is HBM-bound on the fp32 pre-activations (see traffic)"),
                    "all_model_kernels": kernels}
        # K1 (segmentation) and K4 (decode) are the HBM-bound kernels of SURVEY.md section 8(d)
        n_samp = int(slabs[0].sig_off[-1])
        hbm_kernels = {}
        for k, nbytes in (("read_stats", 2 * n_samp * 2),                       # median pass + MAD pass over int16
                          ("base_features", 2 * n_samp + n_bases * (4 + 1 + 8 + 24)),   # samples, starts, base, ev_mean/std, 6 features
                          ("decode", n_bases * (2 + 1 + 2) + 8 * R)):
            t_ms = stage_ms.get(k, 0.0)
            if t_ms > 0:
                gbs = nbytes * args.steps / (t_ms * 1e-3) / 1e9
                hbm_kernels[k] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                                  "bytes_per_step": nbytes, "ms_per_step": t_ms / args.steps}
        roofline["hbm_kernels"] = hbm_kernels
        whole = {"achieved_tflops": FLOP_PER_BASE * value / world / 1e12, "frac_of_bf16_sustained":
                 FLOP_PER_BASE * value / world / 1e12 / peak}
        cpu = None
        if not args.no_cpu_baseline and world == 1:        # rank 0 at N = 1 only
            nreads = max(1, args.cpu_sample_bases // READ_LEN)
            v, dt, nb = cpu_oracle_bases_per_sec(nreads, READ_LEN, False)
            cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": "%d synthetic cfg2 read(s) x %d bases (%.1f s), batched oracle (CNN once per base)" % (nreads, READ_LEN, dt)}
            # the oracle as the checker: the CUDA path must give the same revised bytes on the read the CPU just timed
            sb, want = cpu_oracle_bases_per_sec.last
            got = rv.revise_batch(engine.split_batch(sb, [0])).sequence(0)
            cpu["gpu_matches_oracle_on_sample"] = bool(got == want)
            if got != want:
                raise SystemExit("bench.py: CUDA result differs from the CPU oracle on the sampled read -- number withheld")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "cfg2: %s model, synthetic 10 kb reads (4 kHz signal), slab of %d reads "
                                       "(%d bases) per step per GPU" % (args.species, R, n_bases),
                           "reads_per_step": R, "bases_per_step": n_bases, "windows_per_step": n_win,
                           "l2": "two alternating slabs; per-step activations (%.1f GB) exceed the 126 MB L2" %
                                 (n_win * 11 * 544 * 4 / 1e9)},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": slabs[0].h2d_bytes(),
                        "d2h_bytes_per_step": d2h // max(args.steps, 1)},
                "gpu_launches": int(gpu_launches), "clocks": clocks, "roofline": roofline, "whole_path": whole,
                "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
                "stage_note": ("lstm0 (read_rnn1) and heads (heads tail) are launched on a low-priority side stream: read_rnn1 of the next "
                               "chunk runs under rec2 (fused total_rnn1, 128 of 148 SMs), the heads tail under the next chunk's persistent "
                               "kernels; their times overlap other stages, the stage times do not add up to ms_per_step") if overlapped else None,
                "device_vs_host_path_identical": same}
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    rv.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
