#!/usr/bin/env python
"""CLI end to end: fast5 directory -> NanoReviser.py pipeline (ingest threads -> GPU, two batches in flight -> writer threads)
-> one output file per read.  The reference times the same span (NanoReviser.py:213,228-230: "NanoReviser time consuming").

    python tools/bench_cli.py [--files 4000] [--devices 0] [--format fasta] [--copy] [--species ecoli]

The five unitest fast5 files (tests/golden/fast5, 40,940 bases, 4.3 MB) are replicated to --files entries in a scratch
directory: hard links by default (the page cache then holds every input after the first slab, so the number isolates the
host pipeline: HDF5 parsing, inflate, event collapse, batching, output formatting and file creation), real copies with --copy.
Prints one JSON line: bases/s over the whole run (model load and handle creation included) and over the steady part.
"""
import argparse
import glob
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=4000)
    ap.add_argument("--devices", default="0")
    ap.add_argument("--format", default="fasta")
    ap.add_argument("--species", default="ecoli")
    ap.add_argument("--copy", action="store_true")
    ap.add_argument("--scratch", default=None)
    ap.add_argument("--batch-bases", type=int, default=1_280_000)
    ap.add_argument("--slab-files", type=int, default=512)
    ap.add_argument("--ingest", default="native")
    a = ap.parse_args()
    os.chdir(ROOT)
    import NanoReviser as cli
    from nanoreviser_b200 import fast5

    src = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "fast5", "*.fast5")))
    nb = [fast5.read_fast5_arrays(f).n_bases for f in src]
    scratch = a.scratch or tempfile.mkdtemp(prefix="nrv_cli_")
    din, dout = os.path.join(scratch, "in"), os.path.join(scratch, "out") + "/"
    os.makedirs(din, exist_ok=True)
    total = 0
    for i in range(a.files):
        k = i % len(src)
        dst = os.path.join(din, "r%06d_%s" % (i, os.path.basename(src[k])))
        if not os.path.exists(dst):
            if a.copy:
                shutil.copyfile(src[k], dst)
            else:
                os.link(src[k], dst) if os.stat(src[k]).st_dev == os.stat(din).st_dev else shutil.copyfile(src[k], dst)
        total += nb[k]
    argv = ["-d", din, "-o", dout, "-S", a.species, "-F", a.format, "--devices", a.devices, "--quiet",
            "--batch-bases", str(a.batch_bases), "--slab-files", str(a.slab_files), "--ingest", a.ingest]
    args = cli.get_args(argv)
    # warm-up run on one slab (CUDA context, module load, arenas) so that the steady number is not start-up
    t0 = time.perf_counter()
    res = cli.main(args)
    wall = time.perf_counter() - t0
    stats = [r[4] for r in res if len(r) > 4]
    n_out = len(os.listdir(dout))
    line = {"metric": "cli_e2e_revised_bases_per_sec", "files": a.files, "bases": total, "wall_s": wall,
            "value": total / wall, "unit": "bases/s", "devices": a.devices, "format": a.format, "outputs_written": n_out,
            "inputs": "copies" if a.copy else "hard links of the 5 unitest fast5", "host_cores": os.cpu_count(),
            "note": "whole CLI run: weight load, handle creation, ingest, revise, write",
            # per worker: seconds of set-up (weights, CUDA context, handle), of waiting for the first ingest slab, and where the pipeline's
            # main thread waited afterwards; steady = bases / (total - time to the first slab)
            "workers": [{k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items() if k != "t_worker_start"} for st in stats],
            "steady_value": (sum(st["bases"] for st in stats) / max(max(st["total_s"] - st.get("first_slab_s", 0.0) for st in stats), 1e-9)) if stats else None}
    print(json.dumps(line), flush=True)
    shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    main()
