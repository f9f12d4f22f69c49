"""CPU: input variants of SURVEY.md section 8(f) ranks 1 and 3 -- VBZ-compressed signals (HDF5 filter 32020), legacy
Albacore <= 0.0 event tables and multi-read containers -- through both product readers (nanoreviser_b200/fast5.py on h5mini,
and the native nrv_ingest_fast5).  The fixtures under tests/golden/fast5_variants/ are synthesised from a real fixture read by
tests/golden/make_variant_fixtures.py; the VBZ chunk payloads and the decoder unit vectors come out of the reference's own
plugin binary (nanorevutils/utils/lib/libvbz_hdf_plugin.so), so the decoders are pinned against the reference's bytes."""
import glob
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VAR = os.path.join(ROOT, "tests", "golden", "fast5_variants")
FIELDS = ("a0", "starts", "length", "bases", "signal", "ev_mean", "ev_std")


def _truncated_fixture(k, n_events=2500):
    """what make_variant_fixtures.py cut out of fixture k, read from the ORIGINAL file with the reference-pinned oracle"""
    from nanoreviser_b200 import fast5, h5mini
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "fast5", "*.fast5")), key=os.path.getsize)
    with h5mini.File(files[k]) as f:
        ev = f["/Analyses/Basecall_1D_000/BaseCalled_template/Events"][()][:n_events]
        sig = list(f["/Raw/Reads"].values())[0]["Signal"][()][:int(ev["start"][-1] + ev["length"][-1]) + 37]
    start, bases, m, s = fast5.collapse_events(ev["start"], ev["mean"], ev["stdv"], ev["model_state"], ev["move"])
    return dict(a0=int(start[0]), starts=start - start[0], bases=bases, signal=sig, ev_mean=m, ev_std=s)


def test_vbz_decoder_against_the_plugins_own_chunks():
    from nanoreviser_b200 import h5mini
    v = np.load(os.path.join(VAR, "vbz_chunks.npz"))
    names = sorted({k.rsplit("_", 1)[0] for k in v.files})
    assert len(names) == 20
    for nm in names:
        assert h5mini.vbz_decompress(v[nm + "_comp"].tobytes(), v[nm + "_cd"]) == v[nm + "_raw"].tobytes(), nm
    with pytest.raises(h5mini.H5Error):
        h5mini.vbz_decompress(v["real_v0z1_comp"].tobytes()[:40], v["real_v0z1_cd"])          # truncated zstd frame


@pytest.mark.parametrize("name", ["plain", "vbz_v0", "vbz_v1", "vbz_nozstd", "legacy"])
def test_python_reader_on_variants(name):
    from nanoreviser_b200 import fast5
    from oracle import nanorev_oracle as orc
    want = _truncated_fixture(0)
    fn = os.path.join(VAR, name + ".fast5")
    r = fast5.read_fast5_arrays(fn)
    for k in ("a0", "starts", "bases", "signal", "ev_mean", "ev_std"):
        assert np.array_equal(getattr(r, k), want[k]), k
    # and the reference-pinned oracle (restatement of get_read_data incl. the legacy rescaling, :65-75) sees the same read
    a0, starts, length, bases, signal, m, s = orc.get_read_data(fn)
    assert a0 == r.a0 and np.array_equal(starts, r.starts) and np.array_equal(length, r.length)
    assert "".join(bases).encode() == r.bases.tobytes() and np.array_equal(signal, r.signal)


def test_native_reader_on_variants():
    from nanoreviser_b200 import engine, fast5
    names = ["plain", "vbz_v0", "vbz_v1", "vbz_nozstd", "legacy", "multi"]
    paths = [os.path.join(VAR, n + ".fast5") for n in names]
    batch, fstat, read_file, a0, members = engine.ingest_fast5(paths, with_names=True, threads=3)
    assert fstat.tolist() == [0] * 6
    assert read_file.tolist() == [0, 1, 2, 3, 4, 5, 5]
    assert members[:5] == [""] * 5 and members[5].startswith("read_00000001") and members[6].startswith("read_00000002")
    assert fast5.list_members(paths[5]) == members[5:] and fast5.list_members(paths[0]) == []
    for i in range(batch.n_reads):
        want = _truncated_fixture(1 if i == 6 else 0)
        s0, s1 = int(batch.sig_off[i]), int(batch.sig_off[i + 1])
        b0, b1 = int(batch.base_off[i]), int(batch.base_off[i + 1])
        assert int(a0[i]) == want["a0"]
        assert np.array_equal(batch.signal[s0:s1], want["signal"][want["a0"]:])
        assert np.array_equal(batch.starts[b0:b1], want["starts"])
        assert np.array_equal(batch.bases[b0:b1], want["bases"])
        assert np.array_equal(batch.ev_mean[b0:b1], want["ev_mean"]) and np.array_equal(batch.ev_std[b0:b1], want["ev_std"])
        # the Python reader agrees member by member
        r = fast5.read_fast5_arrays(paths[int(read_file[i])], member=members[i] or None)
        assert r.a0 == int(a0[i]) and np.array_equal(r.starts, batch.starts[b0:b1]) and np.array_equal(r.signal[r.a0:], batch.signal[s0:s1])


def test_legacy_fixture_detects_a_missing_rescale(tmp_path):
    """the legacy fixture is sensitive: without `start * 4000 - start_time` the collapse gives other starts (so a reader that
    skipped the branch would fail the tests above)"""
    from nanoreviser_b200 import h5mini
    with h5mini.File(os.path.join(VAR, "legacy.fast5")) as f:
        ev = f["/Analyses/Basecall_1D_000/BaseCalled_template/Events"][()]
        st = int(list(f["/Raw/Reads"].values())[0].attrs["start_time"])
        assert "version" not in f["/Analyses/Basecall_1D_000"].attrs
    assert ev["start"].dtype == np.float64 and st > 0
    assert not np.array_equal(np.trunc(ev["start"]).astype(np.int64), np.trunc(ev["start"] * 4000 - st).astype(np.int64))


def test_native_reader_rejects_corrupt_vbz(tmp_path):
    """a damaged VBZ payload is a per-file CORRUPT status, never a crash"""
    from nanoreviser_b200 import engine
    raw = bytearray(open(os.path.join(VAR, "vbz_v0.fast5"), "rb").read())
    k = raw.find(b"\x28\xb5\x2f\xfd")              # first zstd frame
    assert k > 0
    for flip in (k + 1, k + 6, k + 20, k + 100, k - 4, k - 3):       # zstd magic / header / payload, the size prefix
        bad = bytearray(raw)
        bad[flip] ^= 0xFF
        fn = tmp_path / ("bad_%d.fast5" % flip)
        fn.write_bytes(bytes(bad))
        _b, fstat, _rf, _a0 = engine.ingest_fast5([str(fn)])
        assert int(fstat[0]) == engine.INGEST_CORRUPT
