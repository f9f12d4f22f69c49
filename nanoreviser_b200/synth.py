"""Synthetic reads of the shapes BASELINE.json names; host-side, seeded.

cfg2  "ecoli model, synthetic 100k reads x 10 kb (4 kHz signal)":  N = 10,000 bases per read, per-base
      duration ``5 * (1 + Geom(p=0.5625))`` samples (4 kHz / 450 b/s nominal model, mean 8.89),
      ``last_dur`` in {3, 5}; raw ``int16 = round(shift_r + scale_r * (level[3-mer] + N(0,1)))`` with
      ``shift_r ~ U{365..810}``, ``scale_r ~ U{20..76}`` clipped to [-605, 1805] (fixture range);
      ``ev_mean ~ N(104.7, 19.3)`` clipped [37, 182]; ``ev_std ~ |N(0, 7.1)| + 0.5``.
cfg3  human model, read lengths ``LogNormal(mu=9.0935, sigma=0.9)`` clipped to [500, 300000].
cfg5  long-read stress: lengths uniform in [100 kb, 300 kb].

Every read is generated from ``(seed, read_id)`` alone, so any shard of the job can be produced
independently on any rank.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from .engine import Batch

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_LEVELS = np.random.default_rng(0xC0FFEE).normal(0.0, 1.0, size=64)


def read_lengths(cfg: str, n_reads: int, seed: int = 0x5EED, first_id: int = 0) -> np.ndarray:
    ids = np.arange(first_id, first_id + n_reads)
    if cfg == "cfg2":
        return np.full(n_reads, 10_000, dtype=np.int64)
    out = np.empty(n_reads, dtype=np.int64)
    for k, i in enumerate(ids):
        rng = np.random.default_rng([seed, int(i), 1])
        if cfg == "cfg3":
            out[k] = int(np.clip(rng.lognormal(9.0935, 0.9), 500, 300_000))
        elif cfg == "cfg5":
            out[k] = int(rng.integers(100_000, 300_001))
        else:
            raise ValueError(cfg)
    return out


def make_read(n_bases: int, seed: int, read_id: int):
    """-> (signal int16[S], starts int32[N], bases uint8[N], ev_mean f32[N], ev_std f32[N], last_dur)."""
    rng = np.random.default_rng([seed ^ 0x5EED, int(read_id)])
    N = int(n_bases)
    b = rng.integers(0, 4, size=N, dtype=np.int64)
    bases = _ACGT[b]
    dur = 5 * rng.geometric(0.5625, size=N).astype(np.int64)        # 5 * (1 + Geom0)
    last_dur = 3 if rng.random() < 0.5 else 5
    dur[-1] = last_dur
    starts = np.zeros(N, dtype=np.int64)
    np.cumsum(dur[:-1], out=starts[1:])
    tail = int(rng.integers(0, 200))                                 # raw signal continues after the last event
    S = int(starts[-1] + last_dur + tail)
    ctx = (np.roll(b, 1) * 16 + b * 4 + np.roll(b, -1)) & 63
    level = _LEVELS[ctx]
    per_sample = np.repeat(level, dur)
    per_sample = np.concatenate([per_sample, np.full(S - per_sample.shape[0], 0.0)])
    shift_r = float(rng.integers(365, 811))
    scale_r = float(rng.integers(20, 77))
    raw = np.rint(shift_r + scale_r * (per_sample + rng.standard_normal(S)))
    signal = np.clip(raw, -605, 1805).astype(np.int16)
    ev_mean = np.clip(rng.normal(104.7, 19.3, size=N), 37, 182).astype(np.float32)
    ev_std = (np.abs(rng.normal(0, 7.1, size=N)) + 0.5).astype(np.float32)
    return signal, starts.astype(np.int32), bases, ev_mean, ev_std, last_dur


def make_batch(lengths: Sequence[int], seed: int = 0, first_id: int = 0, ids: Sequence[int] = None) -> Batch:
    """``ids``: explicit read ids (a shard of a larger job); default ``first_id + i``."""
    R = len(lengths)
    parts = [make_read(int(n), seed, int(ids[i]) if ids is not None else first_id + i) for i, n in enumerate(lengths)]
    sig_off = np.zeros(R + 1, dtype=np.int64)
    base_off = np.zeros(R + 1, dtype=np.int64)
    for i, p in enumerate(parts):
        sig_off[i + 1] = sig_off[i] + p[0].shape[0]
        base_off[i + 1] = base_off[i] + p[1].shape[0]
    cat = lambda k, dt: (np.concatenate([p[k] for p in parts]).astype(dt, copy=False) if R else np.zeros(0, dt))
    return Batch(signal=cat(0, np.int16), sig_off=sig_off, starts=cat(1, np.int32), base_off=base_off,
                 bases=cat(2, np.uint8), ev_mean=cat(3, np.float32), ev_std=cat(4, np.float32),
                 last_dur=np.array([p[5] for p in parts], dtype=np.int32))


from .engine import split_batch  # noqa: E402,F401  (moved to engine.py; kept importable from here for the tests)
