#!/usr/bin/env python
"""Measured precision of the GPU path on the unitest set (81,770 windows, both species) against the fp64 oracle goldens, for the
environment given on the command line (e.g. NRV_F8=0).  Prints one markdown row per species.  usage: gpu_precision.py [LABEL]"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanoreviser_b200 import api, engine, fast5, weights  # noqa: E402

label = sys.argv[1] if len(sys.argv) > 1 else "default"
reads = [fast5.read_fast5_arrays(f) for f in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "fast5", "*.fast5")))]
for sp in ("ecoli", "human"):
    m1, m2 = weights.load_species(sp, os.path.join(ROOT, "model"))
    gold = np.load(os.path.join(ROOT, "tests", "golden", "forward_%s.npz" % sp))
    with engine.Reviser(m1, m2, device=0) as rv:
        out = api.revise_reads(reads, reviser=rv, want_labels=True, want_probs=True)
    w0 = 0
    d = [0.0, 0.0]
    flips = [0, 0]
    worst_margin = 1.0
    tot = seq = 0
    for k, r in enumerate(reads):
        M = r.n_bases - 11
        for mi, (P, y) in enumerate(((out.p1[w0:w0 + M], out.y1[w0:w0 + M]), (out.p2[w0:w0 + M], out.y2[w0:w0 + M]))):
            G = gold["r%d_P%d_f64" % (k, mi + 1)]
            d[mi] = max(d[mi], float(np.abs(P - G).max()))
            diff = np.nonzero(y != gold["r%d_y%d_f64" % (k, mi + 1)])[0]
            flips[mi] += len(diff)
            if len(diff):
                s = np.sort(G[diff], 1)
                worst_margin = min(worst_margin, float((s[:, -1] - s[:, -2]).max()))
        tot += M
        seq += int(out.sequence(k) == gold["r%d_revised" % k].tobytes().decode())
        w0 += M
    print("| %s | %s | %.2e | %.2e | %d + %d of %d%s | %d / 5 |" % (
        label, sp, d[0], d[1], flips[0], flips[1], tot, (" (fp64 margin of the flipped windows <= %.1e)" % worst_margin) if sum(flips) else "", seq))
