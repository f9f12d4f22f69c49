#!/usr/bin/env python
"""fast5 ingest rate: native C++ reader (nrv_ingest_fast5) at 1..N host threads vs the
Python reader, on copies of the 5 unitest fast5 (page-cache resident).  Prints a markdown table.
  python tools/bench_ingest.py [copies=40] > profiles/r01_ingest.md"""
import glob, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanoreviser_b200 import engine, fast5  # noqa: E402

copies = int(sys.argv[1]) if len(sys.argv) > 1 else 40
src = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "fast5", "*.fast5")))
tmp = tempfile.mkdtemp(prefix="nrv_ingest_")
paths = []
for c in range(copies):
    for f in src:
        p = os.path.join(tmp, "c%03d_%s" % (c, os.path.basename(f)))
        shutil.copy(f, p)
        paths.append(p)
for p in paths:
    open(p, "rb").read()          # page cache
print("# fast5 ingest: %d single-read files (%d copies of the unitest set), host cores: %d\n" % (len(paths), copies, os.cpu_count()))
print("| reader | threads | s | files/s | M bases/s | MB/s inflated int16 |\n|---|---|---|---|---|---|")
t0 = time.perf_counter()
reads = [fast5.read_fast5_arrays(p) for p in paths[:len(paths) // 4]]
tp = time.perf_counter() - t0
nb = sum(r.n_bases for r in reads); ns = sum(len(r.signal) for r in reads)
print("| python (h5mini + numpy) | 1 | %.3f | %.0f | %.2f | %.0f |" % (tp, len(reads) / tp, nb / tp / 1e6, ns * 2 / tp / 1e6))
th = 1
while th <= (os.cpu_count() or 1):
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        batch, st, rf, a0 = engine.ingest_fast5(paths, threads=th)
        t = time.perf_counter() - t0
        best = t if best is None else min(best, t)
    assert st.tolist() == [0] * len(paths)
    print("| native C++ | %d | %.3f | %.0f | %.2f | %.0f |" % (th, best, len(paths) / best, batch.n_bases / best / 1e6, int(batch.sig_off[-1]) * 2 / best / 1e6))
    th *= 2
shutil.rmtree(tmp)
