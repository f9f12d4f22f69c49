"""CPU: host logic of the product (HDF5 reader, ingest, weights, batching, writers, CLI flags,
work queue) and the C-ABI surface (library loads and exports every symbol include/nrv.h declares)."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_h5mini_structure(fast5_files):
    from nanoreviser_b200 import h5mini
    for fn in fast5_files:
        with h5mini.File(fn) as f:
            assert set(f.keys()) >= {"Analyses", "Raw"}
            g = f["/Analyses/Basecall_1D_000"]
            assert g.attrs["version"] == b"2.0.2"
            ev = f["/Analyses/Basecall_1D_000/BaseCalled_template/Events"][()]
            assert ev.dtype.itemsize == 41 and ev.dtype.names[:4] == ("mean", "start", "stdv", "length")
            assert set(np.unique(ev["move"])) <= {0, 1, 2} and np.all(ev["length"] == 5)
            summ = f["/Analyses/Basecall_1D_000/Summary/basecall_1d_template"].attrs
            assert int(summ["num_events"]) == len(ev)
            rd = list(f["/Raw/Reads"].values())[0]
            sig = rd["Signal"][()]
            assert sig.dtype == np.int16 and len(sig) == int(rd.attrs["duration"])
            with pytest.raises(KeyError):
                f["/Analyses/Nope"]
    with pytest.raises(h5mini.H5Error):
        h5mini.File(__file__)


def test_weights_loader_all_files():
    from nanoreviser_b200 import weights
    for sp in ("ecoli", "human"):
        m1, m2 = weights.load_species(sp, os.path.join(ROOT, "model"))
        assert (m1.window, m1.n_class, m2.window, m2.n_class) == (11, 6, 11, 5)      # F3: W = 11, not 13
        assert m1.n_params() == 595_300 and m2.n_params() == 595_283
        assert m1.lstm[2][0].kernel.shape == (192, 512) and m1.lstm[2][1].recurrent.shape == (128, 512)
    assert weights.model_paths("ecoli") == ("./model/ecoli/ecoli_win13_50ep_model1.h5",
                                            "./model/ecoli/ecoli_win13_50ep_model2.h5")


def test_product_ingest_matches_reference_outputs(fast5_files, seg_golden):
    from nanoreviser_b200 import fast5
    for k, fn in enumerate(fast5_files):
        a0, starts, length, bases, signal, em, es = fast5.get_read_data(fn, "Basecall_1D_000", "BaseCalled_template")
        g = lambda n: seg_golden["r%d_%s" % (k, n)]
        assert a0 == int(g("a0")) and np.array_equal(starts, g("starts")) and np.array_equal(length, g("length"))
        assert "".join(bases).encode() == g("bases").tobytes()
        assert np.array_equal(np.asarray(em), g("ev_mean")) and np.array_equal(np.asarray(es), g("ev_std"))
        seq, qul = fast5.extract_fastq(fn)
        assert "".join(bases)[5:-5] == seq and len(seq) == len(qul)      # Fastq[7:-7] vs bases == Fastq[2:-2]
    with pytest.raises(NotImplementedError):
        fast5.get_read_data(__file__, "Basecall_1D_000", "BaseCalled_template")
    with pytest.raises(RuntimeError):
        fast5.get_read_data(fast5_files[0], "Basecall_1D_999", "BaseCalled_template")


def test_collapse_events_moves():
    from nanoreviser_b200 import fast5
    st = np.array([100, 105, 110, 115, 120], np.uint64)
    mv = np.array([1, 0, 2, 1, 3], np.int32)
    ms = np.array([b"AACGT", b"ACGTA", b"CGTAC", b"GTACG", b"TACGA"], dtype="S5")
    mean = np.arange(5, dtype=np.float32); sd = mean + 10
    start, bases, m, s = fast5.collapse_events(st, mean, sd, ms, mv)
    assert start.tolist() == [100, 110, 112, 115, 120]
    assert bases.tobytes() == b"CGTAC"          # move2: state[1] then state[2]; move 3 treated as 1
    assert m.tolist() == [0, 2, 2, 3, 4] and s.tolist() == [10, 12, 12, 13, 14]


def test_pack_batch_and_synth(reads):
    from nanoreviser_b200 import engine, synth
    b = engine.pack_batch(reads)
    assert b.n_reads == 5 and b.n_bases == 40_940 and b.n_windows(11) == 40_940 - 55
    assert b.signal.dtype == np.int16 and b.starts.dtype == np.int32
    for i, r in enumerate(reads):
        assert np.array_equal(b.signal[b.sig_off[i]:b.sig_off[i + 1]], r.signal[r.a0:])
        assert b.starts[b.base_off[i]] == 0 and b.last_dur[i] in (3, 5)
    s1 = synth.make_batch([10_000, 500], seed=3)
    s2 = synth.make_batch([10_000, 500], seed=3)
    assert np.array_equal(s1.signal, s2.signal) and np.array_equal(s1.bases, s2.bases)
    assert s1.n_bases == 10_500 and abs(s1.sig_off[1] / 10_000 - 8.89) < 0.3       # 4 kHz / 450 b/s
    assert s1.signal.min() >= -605 and s1.signal.max() <= 1805
    sub = synth.split_batch(s1, [1])
    assert sub.n_bases == 500 and np.array_equal(sub.bases, s1.bases[10_000:])
    L = synth.read_lengths("cfg3", 2000)
    assert L.min() >= 500 and L.max() <= 300_000 and 8_300 < np.median(L) < 9_500   # exp(9.0935) = 8.9 k


def test_writers_and_names(tmp_path):
    from nanoreviser_b200 import api
    fn = str(tmp_path / "o.fasta")
    api.prep_read_fasta("/x/y/a b.fast5", fn, list("ACGT"))
    assert open(fn).read() == ">a|||b.fast5\nACGT"                       # no trailing newline
    api.prep_read_fastq("/x/y/a b.fast5", fn, list("ACGT"), list("!!!!"))
    assert open(fn).read() == "@a|||b.fast5\nACGT+\n!!!!"               # no newline before '+'
    assert api.out_filename("./out/", "r1.strand.fast5", "fasta") == "./out/r1_out.fasta"
    assert api.get_base_color("G") == 180 and api.get_base_color("N") == 0 and api.get_base_label("-") == 1
    with pytest.raises(NotImplementedError):
        api.prep_read_fasta("a.fast5", str(tmp_path / "nodir" / "x"), list("A"))


def test_cli_flags_match_reference():
    sys.path.insert(0, ROOT)
    import NanoReviser as cli
    a = cli.get_args(["-d", "in/", "-o", "out/"])
    assert (a.species, a.output_format, a.thread, a.basecall_group, a.basecall_subgroup) == \
        ("human", "fasta", 100, "Basecall_1D_000", "BaseCalled_template")
    a = cli.get_args(["-d", "in/", "-o", "out/", "-S", "ecoli", "-F", "fastq", "--thread", "4", "-t", "tmp/",
                      "-e", "f.txt", "-g", "G", "-s", "S", "--test_mode", "--devices", "0,1"])
    assert (a.species, a.output_format, a.thread, a.test_mode, a.devices) == ("ecoli", "fastq", 4, True, "0,1")
    with pytest.raises(SystemExit):
        cli.get_args([])


def test_workqueue_partition_and_batches():
    from nanoreviser_b200 import synth, workqueue
    L = synth.read_lengths("cfg3", 4000).tolist()
    for world in (1, 2, 4, 8):
        parts = workqueue.lpt_partition(L, world)
        assert sorted(i for p in parts for i in p) == list(range(len(L)))        # disjoint and complete
        assert workqueue.imbalance(L, parts) < 1.01
        assert parts == workqueue.lpt_partition(L, world)                        # deterministic
    batches = workqueue.make_batches(range(len(L)), L, 2_000_000)
    assert [i for b in batches for i in b] == list(range(len(L)))
    assert all(sum(L[i] for i in b) <= 2_000_000 or len(b) == 1 for b in batches)
    assert workqueue.make_batches([0, 1], [5_000_000, 10], 1000) == [[0], [1]]
    with pytest.raises(ValueError):
        workqueue.lpt_partition(L, 0)


def test_cabi_library_exports_every_declared_symbol():
    """The C-ABI library loads on a CPU box and exports exactly what include/nrv.h declares."""
    import ctypes
    from nanoreviser_b200 import engine
    hdr = open(os.path.join(ROOT, "include", "nrv.h")).read()
    declared = set(re.findall(r"\b(nrv_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(engine.EXPORTS), declared ^ set(engine.EXPORTS)
    if not os.path.exists(engine.LIB_PATH):
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(engine.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    engine.load_library()
    assert b"sm_100a" in engine.load_library().nrv_version()


def test_training_cabi_exports_and_no_cpu_fallback():
    """include/nrv_train.h: every declared operator is exported by the library and bound by train.py; without a GPU the training
    model refuses to exist."""
    import ctypes
    import torch
    from nanoreviser_b200 import engine, train
    hdr = open(os.path.join(ROOT, "include", "nrv_train.h")).read()
    declared = set(re.findall(r"\b(nrvt_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(train.TRAIN_EXPORTS), declared ^ set(train.TRAIN_EXPORTS)
    lib = ctypes.CDLL(engine.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    if not torch.cuda.is_available():
        with pytest.raises(train.TrainError, match="no CUDA device"):
            train.TrainModel(window=11, n_class=6)
    w = train.init_weights(13, 5, seed=1)
    assert w.feat_k.shape == (78, 16) and w.final_k.shape == (16, 5) and w.lstm[2][0].kernel.shape == (192, 512)
    assert np.allclose(w.lstm[1][1].recurrent @ w.lstm[1][1].recurrent.T, np.eye(64), atol=1e-5)        # orthogonal rows
    assert w.lstm[0][0].bias[16:32].tolist() == [1.0] * 16 and w.lstm[0][0].bias[:16].tolist() == [0.0] * 16


def test_saved_weight_file_is_keras_layout(tmp_path, weights_by_species):
    """train.save_predict_weights writes what weights.load_model_weights (and the independent reader of tests/indep_keras.py) read:
    a round trip of the shipped ecoli model reproduces every array bit for bit under the original dataset paths."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import indep_keras
    from nanoreviser_b200 import train, weights
    m1, _ = weights_by_species("ecoli")
    fn = str(tmp_path / "rt.h5")
    train.save_predict_weights(m1, fn)
    back = weights.load_model_weights(fn)
    for k, v in m1.__dict__.items():
        if isinstance(v, np.ndarray):
            assert np.array_equal(v, getattr(back, k)), k
    for l in range(4):
        for d in range(2):
            for f in ("kernel", "recurrent", "bias"):
                assert np.array_equal(getattr(m1.lstm[l][d], f), getattr(back.lstm[l][d], f))
    orig = indep_keras.H5Scan(os.path.join(ROOT, "model", "ecoli", "ecoli_win13_50ep_model1.h5")).walk()
    mine = indep_keras.H5Scan(fn).walk()
    assert set(orig) == set(mine)
    for k in orig:
        assert np.array_equal(orig[k], mine[k]), k


def test_no_cpu_fallback_without_device(weights_by_species):
    """Without a GPU the product fails loudly; it never routes through the oracle."""
    import torch
    from nanoreviser_b200 import api, engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m1, m2 = weights_by_species("ecoli")
    with pytest.raises(engine.NrvError, match="no CUDA device"):
        engine.Reviser(m1, m2)
    api.set_default(None)
    with pytest.raises(engine.NrvError):
        api.signal_segmentation(np.zeros(100, np.int16), np.arange(0, 50, 5), 5)
    for mod in ("engine", "api", "fast5", "weights", "synth", "workqueue", "h5mini", "build"):
        src = open(os.path.join(ROOT, "nanoreviser_b200", mod + ".py")).read()
        assert "import oracle" not in src and "from oracle" not in src, mod


def test_basecall_phred_and_qual_plumbing(fast5_files, reads):
    """Basecaller qualities for the fastq path (D6'): the Fastq dataset lines up with the event-collapsed bases
    (bases == Fastq_seq[2:-2]) on every fixture; Batch.qual travels through pack_batch / split_batch."""
    import copy
    from nanoreviser_b200 import engine, fast5, synth
    rs = []
    for fn, r in zip(fast5_files, reads):
        q = fast5.basecall_phred(fn, r.bases)
        assert q is not None and q.dtype == np.uint8 and q.shape == (r.n_bases,) and q.max() <= 93
        seq, qul = fast5.extract_fastq(fn)                           # the reference's own trimmed view (Fastq[7:-7])
        assert bytes(q[5:-5] + 33).decode() == qul                   # [2:-2] of the call vs [7:-7]: 5 more on each side
        assert fast5.basecall_phred(fn, r.bases[:-1]) is None        # does not line up -> caller falls back to Phred 40
        rr = copy.copy(r)
        rr.qual = q
        rs.append(rr)
    b = engine.pack_batch(rs)
    assert b.qual is not None and b.qual.shape[0] == b.n_bases
    sub = synth.split_batch(b, [3, 1])
    assert np.array_equal(sub.qual, np.concatenate([rs[3].qual, rs[1].qual]))
    assert engine.pack_batch(reads).qual is None                    # all-or-nothing: no qualities unless every read has them


def test_legacy_version_gate_follows_the_reference(fast5_files, tmp_path):
    """SURVEY.md section 8(f) rank 3 (legacy event tables): a file whose Basecall group says version <= 0.0 takes the reference's
    rescaling branch (start*4000 - start_time, fast5_handeler.py:65-72).  The fixture's integer sample starts then wrap around and
    the reference's own length check (:142-143) rejects the read -- in the oracle, in the Python reader, and the native reader
    declines the file (NRV_INGEST_UNSUPPORTED) so that the CLI sends it through the Python reader."""
    from oracle import nanorev_oracle as orc
    from nanoreviser_b200 import engine, fast5, h5mini
    # the version gate itself, against distutils' LooseVersion semantics restated in the oracle
    for v, legacy in (("0.0", True), ("0", True), ("00.00", True), ("0.0.0", False), ("0.0.1", False), ("1.0", False), ("2.0.2", False),
                      ("0.0a", False), (b"0.0", True)):
        assert (orc._loose_version(v) <= orc._loose_version("0.0")) == legacy, v
        assert fast5._version_le_zero(v) == legacy, v
    raw = bytearray(open(fast5_files[0], "rb").read())
    hits = [i for i in range(len(raw) - 5) if raw[i:i + 5] == b"2.0.2"]
    patched = None
    for off in hits:                                   # the vlen string of /Analyses/Basecall_1D_000 attrs['version'] lives in a global heap
        b = bytearray(raw)
        b[off:off + 5] = b"00.00"
        p = str(tmp_path / ("legacy_%d.fast5" % off))
        open(p, "wb").write(b)
        f = h5mini.File(p, "r")
        v = f["/Analyses/Basecall_1D_000"].attrs["version"]
        f.close()
        if bytes(v) == b"00.00":
            patched = p
    assert patched is not None
    with pytest.raises(RuntimeError, match="Signal is shorter than the Events"):
        orc.get_read_data(patched)
    with pytest.raises(RuntimeError, match="Signal is shorter than the Events"):
        fast5.read_fast5_arrays(patched)
    if os.path.exists(engine.LIB_PATH):
        _, st, _, _ = engine.ingest_fast5([patched, fast5_files[0]], "Basecall_1D_000", "BaseCalled_template", 1)
        assert st[0] != engine.INGEST_OK and st[1] == engine.INGEST_OK


def test_ctypes_mirrors_match_the_c_header(tmp_path):
    """include/nrv.h is plain C (a reference maintainer binds it with ctypes / cgo-style FFI): compile it with gcc and compare every
    field offset and struct size of nrv_batch / nrv_result / nrv_model_weights with the ctypes mirrors in engine.py."""
    import ctypes as C
    import shutil
    import subprocess
    from nanoreviser_b200 import engine
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mirrors = {"nrv_batch": engine._Batch, "nrv_result": engine._Result, "nrv_model_weights": engine._ModelWeights,
               "nrv_lstm_dir": engine._LstmDir}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "nrv.h"', '#include "nrv_train.h"   /* plain C as well */', 'int main(void) {',
             '(void)sizeof(&nrvt_adam);   /* declared, not linked */']
    for cname, cls in mirrors.items():
        lines.append('printf("%s.sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0; }']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "abi")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", exe], check=True)
    got = dict(l.split() for l in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.strip().splitlines())
    for cname, cls in mirrors.items():
        assert int(got[cname + ".sizeof"]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_fp32_division_reproduces_the_fp64_quotient_cast_to_fp32():
    """K2's window normalisation (nrv_cnn.cu stage A) divides (x - shift) by scale ONCE in fp32, where the reference divides in
    fp64 (preprocessing.py:119) and Keras casts to fp32.  x is an int16, shift = np.median a multiple of 0.5, scale = the MAD a
    multiple of 0.25: both operands are exact in fp32 and the quotient of two such numbers never comes within the fp64 rounding
    error of an fp32 rounding boundary, so both routes give the same bits.  Checked on 20 M random operand pairs plus the
    exhaustive numerator range for a few denominators."""
    rng = np.random.default_rng(7)
    n = 20_000_000
    num = (rng.integers(-2 * 65535, 2 * 65535 + 1, n) / 2.0)                 # x - shift: half-integers in [-65535, 65535]
    den = (rng.integers(1, 4 * 65535 + 1, n) / 4.0)                          # MAD: quarter-integers in (0, 65535]
    ref = (num / den).astype(np.float32)                                      # fp64 quotient, one cast
    got = num.astype(np.float32) / den.astype(np.float32)                     # one fp32 division of exact operands
    assert num.astype(np.float32).astype(np.float64).tobytes() == num.tobytes()
    assert den.astype(np.float32).astype(np.float64).tobytes() == den.tobytes()
    assert got.dtype == np.float32 and np.array_equal(ref, got)
    allnum = np.arange(-2 * 65535, 2 * 65535 + 1) / 2.0
    for d in (0.25, 0.75, 3.0, 7.25, 13.5, 77.75, 12345.25, 65535.75):
        assert np.array_equal((allnum / d).astype(np.float32), allnum.astype(np.float32) / np.float32(d))
