"""Pin the CPU oracle against the reference itself and (re)generate tests/golden/*.npz.

Runs ONLY in the build container, where the read-only reference checkout is mounted at
/root/reference (it does not exist on the GPU box; nothing in tests/, smoke() or bench.py reads
it at run time).  TEST INFRASTRUCTURE -- never imported by the product.

What is executed from the reference, unchanged:
  * ``nanorevutils.preprocessing`` -- imported as a module (numpy-only imports).
  * ``get_read_data`` (nanorev_fast5_handeler.py:39-150), ``get_base_1`` / ``get_base_2`` /
    ``label_to_base`` / ``prep_read_fasta`` / ``prep_read_fastq`` (output_handeler.py) -- their
    source is AST-extracted and exec'd (the modules cannot be imported: keras / h5py / albacore
    are not installable here); ``h5py`` is shimmed by nanoreviser_b200.h5mini.
What cannot be executed: the Keras graph (A5-A8) -> "parity unpinned", see nanorev_oracle.py.

Usage:  python oracle/pin_against_reference.py [--no-forward] [--species ecoli human]
"""
from __future__ import annotations

import argparse
import ast
import glob
import hashlib
import os
import sys
import tempfile
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get('NANOREV_REFERENCE', '/root/reference')

from nanoreviser_b200 import h5mini, weights as wts   # noqa: E402
from oracle import nanorev_oracle as orc               # noqa: E402


def _extract(src_path, names, extra_globals):
    tree = ast.parse(open(src_path).read())
    body = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            body.append(node)
        elif isinstance(node, ast.Assign) and any(isinstance(t, ast.Name) and t.id in names for t in node.targets):
            body.append(node)
    mod = ast.Module(body=body, type_ignores=[])
    g = dict(extra_globals)
    exec(compile(mod, src_path, 'exec'), g)
    return g


class _LooseVersion:
    """distutils is gone in Python 3.12; the reference only compares dotted numeric versions."""
    def __init__(self, v):
        if isinstance(v, bytes):
            v = v.decode()
        self.v = [int(p) if p.isdigit() else p for p in str(v).split('.')]

    def __le__(self, o):
        return self.v <= o.v


def reference_namespace():
    sys.path.insert(0, REF)
    import nanorevutils.preprocessing as ref_pre
    h5shim = types.SimpleNamespace(File=h5mini.File)
    g_f5 = _extract(os.path.join(REF, 'nanorevutils', 'nanorev_fast5_handeler.py'),
                    {'get_read_data', 'extract_fastq'},
                    {'h5py': h5shim, 'np': np, 'LooseVersion': _LooseVersion})
    g_out = _extract(os.path.join(REF, 'nanorevutils', 'output_handeler.py'),
                     {'get_base_1', 'get_base_2', 'label_to_base', 'prep_read_fasta', 'prep_read_fastq'},
                     {'os': os})
    return ref_pre, g_f5, g_out


def check_segmentation(ref_pre, files, log):
    worst = 0.0
    for fn in files:
        a0, starts, length, bases, signal, em, es = orc.get_read_data(fn)
        sig = signal[int(a0):]
        r = ref_pre.signal_segmentation(sig, starts, int(length[-1]))
        o = orc.signal_segmentation(sig, starts, int(length[-1]))
        assert r[3] == o[3] and r[4] == o[4], (fn, r[3], o[3], r[4], o[4])
        assert np.array_equal(np.asarray(r[0]), o[0]), fn          # windows bit-exact (fp64)
        assert np.array_equal(np.asarray(r[1]), o[1]), fn          # means bit-exact
        d = np.max(np.abs(np.asarray(r[2]) - o[2]))
        assert np.array_equal(np.asarray(r[2]), o[2]), (fn, d)     # same numpy ops -> bit-exact
        worst = max(worst, d)
    log('A2 signal_segmentation: oracle == reference (bit-exact fp64) on %d reads' % len(files))
    for b in 'ACGTN-Dx':
        assert ref_pre.get_base_color(b) == orc.get_base_color(b)
        assert ref_pre.get_base_label(b) == orc.get_base_label(b)


def check_read_data(g_f5, files, log):
    from nanoreviser_b200 import fast5 as prod_f5
    for fn in files:
        r = g_f5['get_read_data'](fn, 'Basecall_1D_000', 'BaseCalled_template')
        o = orc.get_read_data(fn)
        p = prod_f5.get_read_data(fn, 'Basecall_1D_000', 'BaseCalled_template')
        for cand in (o, p):
            assert r[0] == cand[0]
            assert np.array_equal(r[1], cand[1]) and np.array_equal(r[2], cand[2])
            assert list(r[3]) == list(cand[3])
            assert np.array_equal(r[4], cand[4])
            assert np.array_equal(np.array(r[5]), np.array(cand[5])) and np.array_equal(np.array(r[6]), np.array(cand[6]))
        fq = orc.read_fastq_dataset(fn)
        assert ''.join(r[3]) == fq[1][2:-2], fn                     # known answer
        rq = g_f5['extract_fastq'](fn, None)
        pq = prod_f5.extract_fastq(fn, None)
        assert rq == pq
    log('A1 get_read_data: oracle == product == reference source (h5py shimmed) on %d reads; Fastq KAT ok' % len(files))


def check_decode(g_out, log, n_cases=3000):
    rng = np.random.default_rng(1234)
    ref_gb1 = g_out['get_base_1']
    assert g_out['label_to_base'] == orc.label_to_base
    for case in range(n_cases):
        M = int(rng.integers(1, 60))
        bases = list(rng.choice(list('ACGT'), size=M))
        # bias towards interesting combinations
        y1 = rng.integers(0, 6, size=M)
        y2 = rng.integers(0, 5, size=M)
        if case % 3 == 0:
            y2 = np.clip(y1 - 1, 0, 4)
        if case % 7 == 0:
            y1[:] = 1; y2[:] = 0
        r = ref_gb1(bases, y1, y2 + 2)
        o = orc.get_base_1(bases, y1, y2 + 2)
        assert r == o, (case, r, o)
    # writers
    with tempfile.TemporaryDirectory() as td:
        for name in ('a b c.fast5', 'x.fast5'):
            fn = os.path.join(td, 'o.fasta')
            g_out['prep_read_fasta']('/some/dir/' + name, fn, list('ACGT'))
            assert open(fn).read() == orc.fasta_text('/some/dir/' + name, list('ACGT'))
            g_out['prep_read_fastq']('/some/dir/' + name, fn, list('ACGT'), list('!!#$'))
            assert open(fn).read() == orc.fastq_text('/some/dir/' + name, list('ACGT'), list('!!#$'))
    log('A10/A11 get_base_1, writers: oracle == reference source on %d random cases' % n_cases)


def generate_goldens(ref_pre, g_f5, g_out, species_list, do_forward, log):
    files = sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'fast5', '*.fast5')))
    assert len(files) == 5
    rng = np.random.default_rng(7)
    # ---- species-independent part: reference outputs of A1/A2/A3 --------------------
    seg = {}
    for k, fn in enumerate(files):
        a0, starts, length, bases, signal, em, es = g_f5['get_read_data'](fn, 'Basecall_1D_000', 'BaseCalled_template')
        sig = signal[int(a0):]
        win, smean, sstd, shift, scale = ref_pre.signal_segmentation(sig, starts, int(length[-1]))
        N = len(bases)
        x = orc.feature_columns(bases, smean, sstd, shift, scale, length, em, es)
        rows = np.unique(np.concatenate([np.arange(40), np.arange(N - 40, N), rng.integers(0, N, 300)]))
        pre = 'r%d_' % k
        seg[pre + 'name'] = np.array(os.path.basename(fn))
        seg[pre + 'a0'] = np.int64(a0)
        seg[pre + 'starts'] = np.asarray(starts, dtype=np.int64)
        seg[pre + 'length'] = np.asarray(length, dtype=np.float64)
        seg[pre + 'bases'] = np.frombuffer(''.join(bases).encode(), dtype=np.uint8)
        seg[pre + 'ev_mean'] = np.asarray(em, dtype=np.float32)
        seg[pre + 'ev_std'] = np.asarray(es, dtype=np.float32)
        seg[pre + 'shift'] = np.float64(shift)
        seg[pre + 'scale'] = np.float64(scale)
        seg[pre + 'seg_mean'] = np.asarray(smean, dtype=np.float64)
        seg[pre + 'seg_std'] = np.asarray(sstd, dtype=np.float64)
        seg[pre + 'x'] = x
        seg[pre + 'win_rows'] = rows.astype(np.int64)
        seg[pre + 'win'] = np.asarray(win)[rows]
        seg[pre + 'win_md5'] = np.array(hashlib.md5(np.asarray(win, dtype=np.float32).tobytes()).hexdigest())
        seg[pre + 'n_samples'] = np.int64(len(sig))
        log('golden A1-A3 %s N=%d a0=%d S=%d shift=%g scale=%g last_dur=%d md5(bases)=%s' % (
            os.path.basename(fn)[-30:], N, a0, len(sig), shift, scale, int(length[-1]),
            hashlib.md5(''.join(bases).encode()).hexdigest()))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'segmentation.npz'), **seg)
    if not do_forward:
        return
    # ---- per species: oracle forward (fp64 and fp32), labels, revised strings --------
    for sp in species_list:
        m1, m2 = wts.load_species(sp, os.path.join(ROOT, 'model'))
        out = {}
        for k, fn in enumerate(files):
            t0 = time.time()
            pre = 'r%d_' % k
            r64 = orc.revise_fast5(m1, m2, fn, dt=np.float64, want=('probs',))
            r32 = orc.revise_fast5(m1, m2, fn, dt=np.float32, want=('probs',))
            # the decode step through the REFERENCE get_base_1 (verbatim) on the oracle labels
            M = len(r64['bases']) - m1.window
            core = g_out['get_base_1'](r64['bases'][5:5 + M], r64['y1'], r64['y2'] + 2)
            revised = ''.join(r64['bases'][:5]) + core + ''.join(r64['bases'][5 + M:])
            assert revised == r64['revised']
            out[pre + 'P1_f64'] = r64['P1'].astype(np.float32)
            out[pre + 'P2_f64'] = r64['P2'].astype(np.float32)
            out[pre + 'y1_f64'] = r64['y1'].astype(np.uint8)
            out[pre + 'y2_f64'] = r64['y2'].astype(np.uint8)
            out[pre + 'y1_f32'] = r32['y1'].astype(np.uint8)
            out[pre + 'y2_f32'] = r32['y2'].astype(np.uint8)
            out[pre + 'revised'] = np.frombuffer(revised.encode(), dtype=np.uint8)
            out[pre + 'revised_f32'] = np.frombuffer(r32['revised'].encode(), dtype=np.uint8)
            out[pre + 'fasta'] = np.frombuffer(orc.fasta_text(fn, list(revised)).encode(), dtype=np.uint8)
            d1 = np.max(np.abs(r64['P1'] - r32['P1'])); d2 = np.max(np.abs(r64['P2'] - r32['P2']))
            a1 = np.mean(r64['y1'] == r32['y1']); a2 = np.mean(r64['y2'] == r32['y2'])
            log('golden %s %s: fp32-vs-fp64 max|dP| %.2e / %.2e, argmax agree %.5f / %.5f, '
                'y1 hist %s y2 hist %s, len %d -> %d  (%.0fs)' % (
                    sp, os.path.basename(fn)[-30:], d1, d2, a1, a2,
                    np.bincount(r64['y1'], minlength=6).tolist(), np.bincount(r64['y2'], minlength=5).tolist(),
                    len(r64['bases']), len(revised), time.time() - t0))
        np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'forward_%s.npz' % sp), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--no-forward', action='store_true')
    ap.add_argument('--species', nargs='*', default=['ecoli', 'human'])
    ap.add_argument('--log', default=os.path.join(ROOT, 'tests', 'golden', 'PINNING_LOG.txt'))
    args = ap.parse_args()
    lines = []

    def log(s):
        print(s, flush=True)
        lines.append(s)

    ref_pre, g_f5, g_out = reference_namespace()
    all_files = sorted(glob.glob(os.path.join(REF, 'unitest', '*', 'fast5', '*.fast5')))
    check_read_data(g_f5, all_files, log)
    check_segmentation(ref_pre, all_files, log)
    check_decode(g_out, log)
    generate_goldens(ref_pre, g_f5, g_out, args.species, not args.no_forward, log)
    with open(args.log, 'w') as fp:
        fp.write('\n'.join(lines) + '\n')


if __name__ == '__main__':
    main()
