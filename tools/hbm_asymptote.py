"""K1 (read_stats, base_features) and K4 (decode) at growing batch sizes: is the low HBM fraction at the bench shape (23 MB of samples
per batch) launch latency, or is it the kernels?  One nrv_segment / nrv_decode call per size, device times from the stage timers.
usage (GPU box): python tools/hbm_asymptote.py > gpurun_out/<tag>_hbm_asymptote.md"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanoreviser_b200 import engine, synth, weights  # noqa: E402

PEAK = 6452.8
try:
    import json
    PEAK = float(json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", PEAK))
except Exception:
    pass
m1 = weights.load_model_weights("model/ecoli/ecoli_win13_50ep_model1.h5")
m2 = weights.load_model_weights("model/ecoli/ecoli_win13_50ep_model2.h5")
base = synth.make_batch([10_000] * 128, seed=1)
print("| reads x 10 kb | samples MB | read_stats us | GB/s | frac | base_features us | GB/s | frac | decode us | GB/s | frac |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
with engine.Reviser(m1, m2) as rv:
    rv.set_stage_timing(True)
    for mult in [int(v) for v in os.environ.get("HBM_MULTS", "1,4,16,64").split(",")]:
        R = 128 * mult
        sig = np.tile(base.signal, mult)
        so = np.concatenate([[0], np.cumsum(np.tile(np.diff(base.sig_off), mult))]).astype(np.int64)
        bo = np.concatenate([[0], np.cumsum(np.tile(np.diff(base.base_off), mult))]).astype(np.int64)
        b = engine.Batch(signal=sig, sig_off=so, starts=np.tile(base.starts, mult), base_off=bo, bases=np.tile(base.bases, mult),
                         ev_mean=np.tile(base.ev_mean, mult), ev_std=np.tile(base.ev_std, mult), last_dur=np.tile(base.last_dur, mult))
        N = b.n_bases
        rng = np.random.default_rng(3)
        y1 = rng.integers(0, 6, N, dtype=np.uint8)
        y2 = rng.integers(0, 5, N, dtype=np.uint8)
        for rep in range(3):                                   # the third call is the one reported (arenas grown, L2 state as in a run)
            rv.set_stage_timing(True)                          # zeroes the accumulated stage times
            rv.segment(b)
            s = rv.stage_ms()
            rv.set_stage_timing(True)
            rv.decode(b.base_off, b.bases, y1, y2)
            d = rv.stage_ms()
        nsig = sig.shape[0]
        # algorithmic bytes as in bench.py: K1 reads every sample once; the feature kernel reads them again and writes 6 f32 + 2 f64
        # per base (+ the per-base tables it reads); decode reads 3 bytes and writes <= 2 per base
        by_rs = nsig * 2
        by_bf = nsig * 2 + N * (4 + 1 + 4 + 4 + 6 * 4 + 2 * 8)
        by_dc = N * 5
        row = [str(R), "%.1f" % (nsig * 2 / 1e6)]
        for t_ms, by in ((s["read_stats"], by_rs), (s["base_features"], by_bf), (d["decode"], by_dc)):
            gbs = by / (t_ms * 1e-3) / 1e9
            row += ["%.1f" % (t_ms * 1e3), "%.0f" % gbs, "%.3f" % (gbs / PEAK)]
        print("| " + " | ".join(row) + " |", flush=True)
