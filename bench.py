#!/usr/bin/env python
"""bench.py -- revised bases/s of the revision-inference hot path on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg2|cfg3|cfg5]

A *step* is one pass of the whole hot path (K1 segmentation -> K2 CNN -> K3 Bi-LSTM x2 models + heads
-> K4 decode) over one JOB of synthetic reads that is sharded over the N ranks by the host work queue
INSIDE the timed region: ``workqueue.lpt_partition`` (length-balanced shards from the list of read lengths,
identical on every rank, no communication) -> ``workqueue.make_batches`` (ragged batches under a base budget)
-> one library call per batch.  A job holds N x 1.28 M bases, i.e. per-GPU work is fixed ("weak").

  N = 1 (default config cfg2): BASELINE.json configs[1], "ecoli model, synthetic 100k reads x 10 kb (4 kHz
          signal), 1xB200": 128 reads x 10,000 bases per step -- the steady state of the 100k-read job.
  N > 1 (default config cfg3): configs[2], "human model, synthetic 1M reads N50 20 kb, read-sharded over
          2/4/8 B200": lengths ~ LogNormal(9.0935, 0.9) clipped to [500, 300000], ~96 reads per GPU and step.
  --config cfg5: configs[4], ragged 100-300 kb reads, human model (4 x 1.28 M bases per GPU and step so that
          LPT has enough reads to balance).

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


BUDGET = 1_280_000          # bases per GPU and step (= 128 cfg2 reads)
JOBS = 2                    # distinct jobs, alternated, so that no step finds its inputs in L2


def job_config(cfg: str, world: int, reads_per_step: int):
    """-> (species default, bases per rank and step, batch base budget)"""
    if cfg == "cfg2":
        return "ecoli", reads_per_step * READ_LEN, reads_per_step * READ_LEN
    if cfg == "cfg3":
        return "human", BUDGET, int(BUDGET * 1.25)
    if cfg == "cfg5":
        return "human", 4 * BUDGET, int(BUDGET * 1.25)
    raise ValueError(cfg)


def job_lengths(cfg: str, world: int, per_rank_bases: int, job: int):
    """Read lengths of job `job` (a pure function of (cfg, world, job): every rank derives the same list) and the global
    read id of its first read."""
    from nanoreviser_b200 import synth
    first = 1_000_000 * job
    if cfg == "cfg2":
        n = world * per_rank_bases // READ_LEN
        return np.full(n, READ_LEN, dtype=np.int64), first
    target = world * per_rank_bases
    est = {"cfg3": 13_300, "cfg5": 200_000}[cfg]
    L = synth.read_lengths(cfg, int(target / est * 1.3) + 64, first_id=first)
    n = int(np.searchsorted(np.cumsum(L), target)) + 1
    return L[:n].copy(), first


def plan_step(lengths, rank: int, world: int, batch_budget: int):
    """The host work queue of one step: length-balanced shard of this rank, cut into ragged batches (ranges of the shard's
    reads, which are kept in ascending read order so that a batch is a contiguous slice of the shard's CSR arrays)."""
    from nanoreviser_b200 import workqueue
    parts = workqueue.lpt_partition(lengths, world)
    mine = parts[rank]
    pos = {r: k for k, r in enumerate(mine)}
    batches = [(pos[b[0]], pos[b[-1]] + 1) for b in workqueue.make_batches(mine, lengths, batch_budget)]
    return parts, mine, batches


def cpu_oracle_bases_per_sec(n_reads: int, read_len: int, faithful: bool, seed: int = 99, species: str = "ecoli"):
    """Time the CPU oracle on a bounded sample of the same workload.  Returns (bases/s, seconds, bases)."""
    from nanoreviser_b200 import synth, weights
    from oracle import nanorev_oracle as orc
    m1, m2 = weights.load_species(species, os.path.join(ROOT, "model"))
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


def workload_config(cfg: str, species: str, world: int, per_rank: int, batch_budget: int, reads_per_job: int, n_win_rank: int):
    what = {"cfg2": "cfg2: %s model, synthetic 10 kb reads (4 kHz signal)" % species,
            "cfg3": "cfg3: %s model, synthetic reads N50 20 kb (LogNormal(9.0935, 0.9) clipped to [500, 300000]), read-sharded "
                    "over the ranks" % species,
            "cfg5": "cfg5: %s model, long-read stress, ragged 100-300 kb reads, read-sharded over the ranks" % species}[cfg]
    return {"workload": "%s; one job of %d reads (%d bases per GPU) per step, sharded by the host work queue "
                        "(lpt_partition -> make_batches, inside the timed region)" % (what, reads_per_job, per_rank),
            "reads_per_step": int(reads_per_job), "bases_per_step": int(world * per_rank), "bases_per_gpu_per_step": int(per_rank),
            "batch_base_budget": int(batch_budget),
            "l2": "two alternating jobs; per-step activations (%.1f GB per GPU) exceed the 126 MB L2" % (n_win_rank * 11 * 544 * 4 / 1e9)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cfg = args.config or ("cfg2" if world == 1 else "cfg3")
    species_d, per_rank, batch_budget = job_config(cfg, world, args.reads_per_step)
    species = args.species or species_d
    L, _ = job_lengths(cfg, world, per_rank, 0)
    n = max(1, args.ref_reads)
    vals = []
    for _ in range(max(1, args.warmup > 0)):
        cpu_oracle_bases_per_sec(1, 1000, True, species=species)
    for _ in range(args.steps):
        v, dt, nb = cpu_oracle_bases_per_sec(n, args.ref_read_len, True, species=species)
        vals.append((v, dt, nb))
    v = sum(x[2] for x in vals) / sum(x[1] for x in vals)
    sample = "%d synthetic read(s) x %d bases of the step's job per step, %d steps" % (n, args.ref_read_len, args.steps)
    n_win_rank = max(per_rank - 11 * (len(L) // world), 0)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(x[1] for x in vals) / len(vals),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(cfg, species, world, per_rank, batch_budget, len(L), n_win_rank),
            "sample_bases_per_step": n * args.ref_read_len,
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
    ap.add_argument("--config", default=None, choices=["cfg2", "cfg3", "cfg5"],
                    help="default: cfg2 on one GPU, cfg3 (read-sharded) on several")
    ap.add_argument("--reads-per-step", type=int, default=128, help="cfg2: reads per GPU and step")
    ap.add_argument("--species", default=None)
    ap.add_argument("--ref-reads", type=int, default=1)
    ap.add_argument("--ref-read-len", type=int, default=READ_LEN)
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

    from nanoreviser_b200 import engine, synth, weights, workqueue

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    peaks, peak_src = load_peaks()

    cfg = args.config or ("cfg2" if world == 1 else "cfg3")
    species_d, per_rank, batch_budget = job_config(cfg, world, args.reads_per_step)
    species = args.species or species_d
    m1, m2 = weights.load_species(species, os.path.join(ROOT, "model"))
    rv = engine.Reviser(m1, m2, device=local_rank)
    W = rv.window
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.ExternalStream(rv.stream, device=local_rank)

    # ---- the jobs: lengths are the job description (known to every rank); each rank generates only its shard's reads -------
    jobs = []
    for j in range(JOBS):
        L, first = job_lengths(cfg, world, per_rank, j)
        parts, mine, batches = plan_step(L, rank, world, batch_budget)          # also recomputed inside the timed region
        shard = synth.make_batch([int(L[i]) for i in mine], seed=1000 + j, ids=[first + i for i in mine])
        pin = engine.Batch(**{k: torch.from_numpy(np.ascontiguousarray(getattr(shard, k))).pin_memory().numpy() for k in
                              ("signal", "sig_off", "starts", "base_off", "bases", "ev_mean", "ev_std", "last_dur")})
        dten = {k: torch.from_numpy(getattr(shard, k)).to(dev) for k in ("signal", "starts", "bases", "ev_mean", "ev_std", "last_dur")}
        jobs.append({"L": L, "mine": mine, "shard": pin, "dev": dten, "imbalance": workqueue.imbalance(L, parts),
                     "n_batches": len(batches)})
    max_reads = max(j["shard"].n_reads for j in jobs)
    max_bases = max(j["shard"].n_bases for j in jobs)
    cap = 2 * max_bases + max_reads + 16
    d_rev = torch.empty(cap, dtype=torch.uint8, device=dev)
    d_off = torch.empty(max_reads + 1, dtype=torch.int64, device=dev)
    d_status = torch.empty(max_reads, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    esz = {"signal": 2, "starts": 4, "bases": 1, "ev_mean": 4, "ev_std": 4, "last_dur": 4}

    def step_device(i):
        """one step, inputs resident in HBM: work queue on the host, one asynchronous library call per batch"""
        job = jobs[i % JOBS]
        _, mine, batches = plan_step(job["L"], rank, world, batch_budget)
        sh, t = job["shard"], job["dev"]
        nb = 0
        for r0, r1 in batches:
            s0, b0 = int(sh.sig_off[r0]), int(sh.base_off[r0])
            dptr = {"signal": t["signal"].data_ptr() + 2 * s0, "last_dur": t["last_dur"].data_ptr() + 4 * r0}
            for k in ("starts", "bases", "ev_mean", "ev_std"):
                dptr[k] = t[k].data_ptr() + esz[k] * b0
            n = int(sh.base_off[r1]) - b0
            dres = {"revised": d_rev.data_ptr() + 2 * b0 + r0, "out_off": d_off.data_ptr(), "status": d_status.data_ptr() + 4 * r0}
            rv.revise_batch_device(r1 - r0, sh.sig_off[r0:r1 + 1] - s0, sh.base_off[r0:r1 + 1] - b0, dptr, dres, 2 * n + (r1 - r0) + 16)
            nb += n
        return nb

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
    my_bases = 0
    with torch.cuda.stream(stream):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(args.steps):
            my_bases += step_device(i)
        e1.record(stream)
    rv.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    gpu_launches = rv.launch_count - launches0
    stage_ms = rv.stage_ms()
    stage_launches = rv.stage_launches()
    rv.set_stage_timing(False)

    def gather(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world == 1:
            return [float(x)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    rank_ms = gather(ms)
    rank_bases = gather(my_bases)
    ms_max = max(rank_ms)
    total_bases = sum(rank_bases)
    value = total_bases / (ms_max * 1e-3)

    # ---- e2e: host buffers through the public call (submit / wait: two batches in flight; H2D + kernels + D2H inside) -------
    outs = [engine.ReviseResult(torch.empty(cap, dtype=torch.uint8).pin_memory().numpy(),
                                torch.empty(max_reads + 1, dtype=torch.int64).pin_memory().numpy(),
                                torch.empty(max_reads, dtype=torch.int32).pin_memory().numpy()) for _ in range(2)]

    def run_host(steps):
        """-> (bases, h2d bytes, d2h bytes); keeps two batches in flight"""
        pend = []
        nb = h2d = d2h = 0
        k = 0
        for i in range(steps):
            job = jobs[i % JOBS]
            _, mine, batches = plan_step(job["L"], rank, world, batch_budget)
            for r0, r1 in batches:
                sub = engine.slice_batch(job["shard"], r0, r1)
                if len(pend) == 2:
                    o = rv.wait(pend.pop(0))
                    d2h += int(o.out_off[o.n_reads_done]) + 8 * (o.n_reads_done + 1) + 4 * o.n_reads_done
                out = outs[k & 1]; k += 1
                out.n_reads_done = sub.n_reads
                pend.append(rv.submit(sub, out=out))
                nb += sub.n_bases
                h2d += sub.h2d_bytes()
        for p in pend:
            o = rv.wait(p)
            d2h += int(o.out_off[o.n_reads_done]) + 8 * (o.n_reads_done + 1) + 4 * o.n_reads_done
        return nb, h2d, d2h

    run_host(max(1, min(args.warmup, 2)))
    barrier()
    t0 = time.perf_counter()
    e2e_bases, h2d, d2h = run_host(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    e2e_value = sum(gather(e2e_bases)) / max(gather(e2e_s))
    h2d_all, d2h_all = sum(gather(h2d)), sum(gather(d2h))
    # sanity: the device-resident and the host path produce the same bytes (last batch of the last step)
    job = jobs[(args.steps - 1) % JOBS]
    r0, r1 = plan_step(job["L"], rank, world, batch_budget)[2][-1]
    chk = rv.revise_batch(engine.slice_batch(job["shard"], r0, r1))
    torch.cuda.synchronize()
    b0 = int(job["shard"].base_off[r0])
    got_off = d_off.cpu().numpy()[:r1 - r0 + 1]
    got_rev = d_rev.cpu().numpy()[2 * b0 + r0:2 * b0 + r0 + int(chk.out_off[-1])]
    same = bool(np.array_equal(got_off, chk.out_off)) and bool(np.array_equal(got_rev, chk.revised[:chk.out_off[-1]]))

    n_bases = int(round(total_bases / world / args.steps))            # per GPU and step
    n_reads_rank = int(round(sum(j["shard"].n_reads for j in jobs) / JOBS))
    n_win = max(n_bases - W * n_reads_rank, 0)
    n_samp = int(round(sum(int(j["shard"].sig_off[-1]) for j in jobs) / JOBS))
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
                     "rec1": ("lstm_fused_tc64_kernel (read_rnn11 projection+recurrence, tcgen05)" if os.environ.get("NRV_RNN11") == "single" else
                              "lstm_fused_tc64_pp_kernel (read_rnn11 projection+recurrence, tcgen05, two window tiles in flight per CTA)"),
                     "heads_gemm": "gemm_f16x3_kernel<128,FUSE2> (dense head 128->128->32, tcgen05)",
                     "heads": "heads_tail_thread_kernel (dense 32->6, flatten, feature, softmax, argmax; fp32 SIMT)",
                     "lstm0": "read_rnn1_kernel (read_rnn1, fp32 SIMT, one thread per window pair)"}
            if fused1:      # stage proj2 holds only the CNN-feature gather (no MACs); rec2 is the whole layer
                macs.pop("proj2")
                macs["rec2"] = 3_604_480
                f8mode = os.environ.get("NRV_F8", "1")
                names["rec2"] = ("lstm_fused_pair_kernel<192,128> (total_rnn1 projection+recurrence fused, tcgen05 cta_group::2, "
                                 "cluster of 4, h in TMEM, " + ("3-pass split-fp16 projection, 1 fp16 + 1 e4m3 MMA per K-step in the recurrence)"
                                                                if f8mode not in ("0", "2") and fused2 else "3-pass split-fp16)"))
            if fused2:
                macs.pop("proj3")
                macs["rec3"] = 1_802_240
                f8 = os.environ.get("NRV_F8", "1") != "0" and fused1
                names["rec3"] = ("lstm_fused_pair_kernel<256,64> (total_rnn2 projection+recurrence fused, tcgen05 cta_group::2, "
                                 "h in TMEM, " + ("1 fp16 + 1 e4m3 (kind::f8f6f4) MMA per K-step)" if f8 else "3-pass split-fp16)"))
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
                    "note": ("algorithmic FLOPs (1 pass); the kernel spends 3 fp16 MMA passes (2 fp16-equivalents where the corrections run in e4m3) per product to stay fp32-equivalent "
                             "(effective tensor ceiling = peak / 3) and runs under the 1000 W power cap; fused layers move only "
                             "x and h through HBM (see traffic)") if (fused1 or fused2) else
                            ("algorithmic FLOPs (1 pass); the kernel spends 3 fp16 MMA passes per product to stay "
                             "fp32-equivalent, and is HBM-bound on the fp32 pre-activations (see traffic)"),
                    "all_model_kernels": kernels}
        # K1 (segmentation) and K4 (decode) are the HBM-bound kernels of SURVEY.md section 8(d)
        hbm_kernels = {}
        for k, nbytes in (("read_stats", n_samp * 2),                           # ONE pass over the int16 samples (median and MAD from the histogram)
                          ("base_features", 2 * n_samp + n_bases * (4 + 1 + 8 + 24)),   # samples, starts, base, ev_mean/std, 6 features
                          ("decode", n_bases * (2 + 1 + 2) + 8 * n_reads_rank)):
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
            v, dt, nb = cpu_oracle_bases_per_sec(nreads, READ_LEN, False, species=species)
            cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": "%d synthetic read(s) x %d bases, %s weights (%.1f s), batched oracle (CNN once per base)" % (nreads, READ_LEN, species, dt)}
            # the oracle as the checker: the CUDA path must give the same revised bytes on the read the CPU just timed
            sb, want = cpu_oracle_bases_per_sec.last
            got = rv.revise_batch(engine.split_batch(sb, [0])).sequence(0)
            cpu["gpu_matches_oracle_on_sample"] = bool(got == want)
            if got != want:
                raise SystemExit("bench.py: CUDA result differs from the CPU oracle on the sampled read -- number withheld")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f16x3/f32acc", "data": "synthetic",
                "dtype_note": "tensor-core products are three fp16 passes (x_hi.W_hi + x_lo.W_hi + x_hi.W_lo) with fp32 accumulation; in total_rnn2 and "
                              "in the recurrence of total_rnn1 the two correction passes run as one e4m3 product (kind::f8f6f4) on 8-bit copies, "
                              "weights scaled by a power of two per layer; "
                              "segmentation statistics in fp64, indices / decode in integers",
                "config": workload_config(cfg, species, world, per_rank, batch_budget,
                                          int(round(sum(len(j["L"]) for j in jobs) / JOBS)), n_win),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_all // max(args.steps, 1)),
                        "d2h_bytes_per_step": int(d2h_all // max(args.steps, 1)),
                        "api": "Reviser.submit / Reviser.wait (nrv_submit_batch / nrv_wait_batch): pinned host buffers, two batches in flight"},
                "work_queue": {"partition": "lpt_partition + make_batches inside the timed region, every step",
                               "imbalance_max_over_mean": max(j["imbalance"] for j in jobs),
                               "batches_per_rank_per_step": [j["n_batches"] for j in jobs],
                               "rank_ms": rank_ms, "rank_bases": rank_bases,
                               "tail_ms": ms_max - min(rank_ms)},
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
