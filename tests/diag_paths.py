"""Diagnostic (not a test): accuracy of the fp32-SIMT path and the tcgen05 path against the fp64 oracle goldens.
Run on the GPU box:  python tests/diag_paths.py"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(path):
    import glob
    from nanoreviser_b200 import api, engine, fast5, weights
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "fast5", "*.fast5")))
    reads = [fast5.read_fast5_arrays(f) for f in files]
    for sp in ("ecoli", "human"):
        m1, m2 = weights.load_species(sp, os.path.join(ROOT, "model"))
        gold = np.load(os.path.join(ROOT, "tests", "golden", "forward_%s.npz" % sp))
        with engine.Reviser(m1, m2) as rv:
            out = api.revise_reads(reads, reviser=rv, want_labels=True, want_probs=True)
        w0 = 0
        for k, r in enumerate(reads):
            M = r.n_bases - 11
            for name, P, y in (("P1", out.p1[w0:w0 + M], out.y1[w0:w0 + M]), ("P2", out.p2[w0:w0 + M], out.y2[w0:w0 + M])):
                G = gold["r%d_%s_f64" % (k, name)]
                gy = gold["r%d_y%s_f64" % (k, name[1])]
                d = np.abs(P - G).max()
                diff = np.nonzero(y != gy)[0]
                s = np.sort(G, 1)
                mar = s[:, -1] - s[:, -2]
                print("%s %-5s read %d %s: max|dP| %.2e  label diffs %d %s" % (
                    path, sp, k, name, d, len(diff), ["w%d margin %.1e dP %.1e" % (i, mar[i], np.abs(P[i] - G[i]).max()) for i in diff[:4]]))
            same = out.sequence(k) == gold["r%d_revised" % k].tobytes().decode()
            print("%s %-5s read %d revised identical: %s" % (path, sp, k, same))
            w0 += M


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        for p in ("simt", "tc"):
            env = dict(os.environ, NRV_PATH=p)
            subprocess.run([sys.executable, __file__, p], env=env, check=False)
