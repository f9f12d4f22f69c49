"""Native (C++) fast5 ingest -- include/nrv.h: nrv_ingest_fast5 -- against the Python reader (nanoreviser_b200/fast5.py),
which is itself pinned against the reference's own get_read_data source (oracle/pin_against_reference.py, 105 reads).
Host-only: runs without a GPU (the library links the CUDA runtime statically and makes no CUDA call here)."""
import os
import shutil

import numpy as np
import pytest


@pytest.fixture(scope="module")
def lib_built():
    from nanoreviser_b200 import build, engine
    build.build_lib()
    return engine


def _assert_same(batch, k, r):
    s0, s1 = int(batch.sig_off[k]), int(batch.sig_off[k + 1])
    b0, b1 = int(batch.base_off[k]), int(batch.base_off[k + 1])
    assert np.array_equal(batch.signal[s0:s1], r.signal[r.a0:])                 # signal = raw[a0:] (NanoReviser.py:120)
    assert np.array_equal(batch.starts[b0:b1], r.starts.astype(np.int32))
    assert np.array_equal(batch.bases[b0:b1], r.bases)
    assert np.array_equal(batch.ev_mean[b0:b1].view(np.uint32), r.ev_mean.view(np.uint32))   # f4 passthrough, bit-exact
    assert np.array_equal(batch.ev_std[b0:b1].view(np.uint32), r.ev_std.view(np.uint32))
    assert int(batch.last_dur[k]) == r.last_dur


def test_native_ingest_matches_python_reader(lib_built, fast5_files, reads):
    engine = lib_built
    for threads in (1, 4, 0):
        batch, status, read_file, a0 = engine.ingest_fast5(fast5_files, threads=threads)
        assert status.tolist() == [0] * 5 and read_file.tolist() == list(range(5))
        assert batch.n_reads == 5
        for k, r in enumerate(reads):
            assert int(a0[k]) == r.a0
            _assert_same(batch, k, r)
    # the in-file Albacore known answer: bases == Fastq sequence [2:-2]
    from nanoreviser_b200 import h5mini
    for k, fn in enumerate(fast5_files):
        fq = bytes(h5mini.File(fn)["/Analyses/Basecall_1D_000/BaseCalled_template/Fastq"][()]).decode().split("\n")[1]
        b0, b1 = int(batch.base_off[k]), int(batch.base_off[k + 1])
        assert batch.bases[b0:b1].tobytes().decode() == fq[2:-2]


def test_native_ingest_same_batch_as_pack_batch(lib_built, fast5_files, reads):
    engine = lib_built
    batch, *_ = engine.ingest_fast5(fast5_files)
    ref = engine.pack_batch(reads)
    for name in ("signal", "sig_off", "starts", "base_off", "bases", "ev_mean", "ev_std", "last_dur"):
        assert np.array_equal(getattr(batch, name), getattr(ref, name)), name


def test_native_ingest_error_statuses(lib_built, fast5_files, tmp_path):
    """Failures never abort the batch: per-file status, remaining reads packed in order (NanoReviser.py:114-132)."""
    engine = lib_built
    good = fast5_files[0]
    missing = str(tmp_path / "nope.fast5")
    garbage = str(tmp_path / "garbage.fast5")
    open(garbage, "wb").write(b"not an hdf5 file" * 100)
    truncated = str(tmp_path / "truncated.fast5")
    data = open(good, "rb").read()
    open(truncated, "wb").write(data[:len(data) // 3])
    other_group = str(tmp_path / "copy.fast5")
    shutil.copy(good, other_group)
    paths = [good, missing, garbage, truncated, fast5_files[1]]
    batch, status, read_file, _ = engine.ingest_fast5(paths, threads=2)
    assert status[0] == engine.INGEST_OK and status[4] == engine.INGEST_OK
    assert status[1] == engine.INGEST_OPEN_FAILED and status[2] == engine.INGEST_OPEN_FAILED
    assert status[3] != engine.INGEST_OK
    assert read_file.tolist() == [0, 4] and batch.n_reads == 2
    # a basecall group that does not exist -> "No events" (fast5_handeler.py:76-77), like the Python reader
    _, st, _, _ = engine.ingest_fast5([other_group], basecall_group="Basecall_1D_009")
    assert st.tolist() == [engine.INGEST_NO_EVENTS]
    from nanoreviser_b200 import fast5
    with pytest.raises(RuntimeError):
        fast5.read_fast5_arrays(other_group, basecall_group="Basecall_1D_009")
    # empty list
    b, st, rf, a = engine.ingest_fast5([])
    assert b.n_reads == 0 and len(st) == 0


def test_native_ingest_random_corruption_never_crashes(lib_built, fast5_files, tmp_path):
    """Bit rot inside a valid file must yield a status (or a clean read), never a crash / out-of-bounds read."""
    engine = lib_built
    rng = np.random.default_rng(7)
    data = bytearray(open(fast5_files[2], "rb").read())
    paths = []
    for i in range(40):
        d = bytearray(data)
        for _ in range(int(rng.integers(1, 30))):
            p = int(rng.integers(0, len(d)))
            d[p] = int(rng.integers(0, 256))
        fn = str(tmp_path / ("c%d.fast5" % i))
        open(fn, "wb").write(d)
        paths.append(fn)
    batch, status, read_file, _ = engine.ingest_fast5(paths, threads=4)
    assert len(status) == 40 and batch.n_reads == int((status == 0).sum())


def test_native_reader_extracts_basecall_qualities(fast5_files):
    """-F fastq: the native reader returns the basecaller's Phred scores of the event-collapsed bases (Fastq quality - 33 of
    Fastq_seq[2:-2]) exactly as the Python reader does; all-or-nothing per slab."""
    from nanoreviser_b200 import engine, fast5
    batch, st, rf, _ = engine.ingest_fast5(list(fast5_files), "Basecall_1D_000", "BaseCalled_template", 2)
    assert list(st) == [engine.INGEST_OK] * len(fast5_files) and batch.qual is not None and batch.qual.shape[0] == batch.n_bases
    for k, i in enumerate(rf):
        b0, b1 = int(batch.base_off[k]), int(batch.base_off[k + 1])
        want = fast5.basecall_phred(fast5_files[int(i)], batch.bases[b0:b1])
        assert np.array_equal(batch.qual[b0:b1], want)


def _patched(data: bytes, pos: int, new: bytes) -> bytes:
    return data[:pos] + new + data[pos + len(new):]


def test_native_ingest_targeted_corruptions(lib_built, fast5_files, tmp_path):
    """File-supplied sizes and offsets are never trusted (round-1 advisor findings): a record count whose byte size wraps in
    64 bits, compound member offsets beyond the record, and a cyclic chunk B-tree must each yield a per-file status -- the
    worker neither crashes nor hangs, and the other files of the slab are still packed."""
    import struct
    from nanoreviser_b200 import h5mini
    engine = lib_built
    good = fast5_files[0]
    data = open(good, "rb").read()
    n_events = len(h5mini.File(good)["/Analyses/Basecall_1D_000/BaseCalled_template/Events"][()])
    cases = {}
    # (1) Events dataspace: every 8-byte occurrence of the event count -> 2^61 + 1  (E * itemsize wraps to a tiny range)
    d = data
    pos, hits = 0, 0
    needle = struct.pack("<Q", n_events)
    while True:
        pos = d.find(needle, pos)
        if pos < 0:
            break
        d = _patched(d, pos, struct.pack("<Q", (1 << 61) + 1))
        pos += 8
        hits += 1
    assert hits >= 1
    cases["wrap"] = d
    # (2) compound members: offset of `move` / `start` / `model_state` far beyond the record
    for name in (b"move", b"start", b"model_state"):
        key = name + b"\x00" * (8 - len(name) % 8 if len(name) % 8 else 8)
        pos = data.find(key)
        assert pos > 0, name
        cases["member_" + name.decode()] = _patched(data, pos + len(key), struct.pack("<I", 0x7FFFFFF0))
    # (3) chunk B-tree of the signal: make the leaf an internal node whose first child is itself
    pos = 0
    trees = []
    while True:
        pos = data.find(b"TREE", pos)
        if pos < 0:
            break
        if data[pos + 4] == 1:           # node type 1 = raw-data chunks
            trees.append(pos)
        pos += 4
    assert trees
    t = trees[0]
    d = _patched(data, t + 5, b"\x01")                                   # level 1: children are nodes
    keysz = 8 + 8 * 2
    d = _patched(d, t + 24 + keysz, struct.pack("<Q", t))                # first child -> the node itself
    cases["cycle"] = d
    paths = [good]
    for k, v in cases.items():
        fn = str(tmp_path / (k + ".fast5"))
        open(fn, "wb").write(v)
        paths.append(fn)
    paths.append(fast5_files[1])
    batch, status, read_file, _ = engine.ingest_fast5(paths, threads=2)
    assert status[0] == engine.INGEST_OK and status[-1] == engine.INGEST_OK
    assert all(s != engine.INGEST_OK for s in status[1:-1]), dict(zip(["good"] + list(cases) + ["good2"], status.tolist()))
    assert read_file.tolist() == [0, len(paths) - 1]
