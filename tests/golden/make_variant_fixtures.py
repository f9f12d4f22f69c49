"""Generates tests/golden/fast5_variants/ (run in the BUILD container: needs /root/reference for the VBZ plugin binary).

The reference's fixtures are all Albacore 2.0.2 single-read fast5 with a deflate-compressed Signal.  This script re-writes the
smallest of them (its first 2,500 events; tests/h5write.py) as the input variants SURVEY.md section 8(f) ranks 1 and 3 name:

  plain.fast5          the same content through the test writer (proves the writer: both readers must give the fixture's arrays)
  vbz_v0.fast5         Signal chunked with HDF5 filter 32020 (VBZ), cd_values [0, 2, 1, 1]: the chunk payloads are produced by
  vbz_v1.fast5         the reference's OWN plugin binary (nanorevutils/utils/lib/libvbz_hdf_plugin.so: vbz_filter), v1 = [1, 2, 1, 1]
  vbz_nozstd.fast5     cd_values [0, 2, 1, 0]: zig-zag delta + streamvbyte only
  legacy.fast5         an Albacore <= 0.0 table (nanorev_fast5_handeler.py:65-75): no `version` attribute, `start` / `length` as
                       float64 seconds, `start_time` on the raw read
  multi.fast5          a multi-read container: /read_<id>/{Raw/Signal, Analyses/...} for two reads (an extension: the reference
                       itself only reads single-read files, :132-133)

and vbz_chunks.npz: raw / compressed chunk pairs straight from the plugin (random, extreme and real data) for the decoder
unit tests.  Everything is deterministic; the outputs are committed.
"""
import ctypes
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import h5write  # noqa: E402
from nanoreviser_b200 import h5mini  # noqa: E402

PLUGIN = "/root/reference/nanorevutils/utils/lib/libvbz_hdf_plugin.so"
OUT = os.path.join(HERE, "fast5_variants")


def plugin_filter():
    libc = ctypes.CDLL("libc.so.6")
    libc.malloc.restype = ctypes.c_void_p
    libc.malloc.argtypes = [ctypes.c_size_t]
    lib = ctypes.CDLL(PLUGIN)
    flt = getattr(lib, "_Z10vbz_filterjmPKjmPmPPv")     # size_t vbz_filter(flags, cd_nelmts, cd_values, nbytes, *buf_size, **buf)
    flt.restype = ctypes.c_size_t
    flt.argtypes = [ctypes.c_uint, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t),
                    ctypes.POINTER(ctypes.c_void_p)]

    def run(data: bytes, cd, reverse=False) -> bytes:
        p = libc.malloc(max(len(data), 1))
        ctypes.memmove(p, data, len(data))
        buf, bs = ctypes.c_void_p(p), ctypes.c_size_t(len(data))
        cdv = (ctypes.c_uint * len(cd))(*cd)
        r = flt(0x100 if reverse else 0, len(cd), cdv, len(data), ctypes.byref(bs), ctypes.byref(buf))
        assert r > 0, "vbz_filter failed"
        return ctypes.string_at(buf.value, r)

    return run


def load_fixture(fn, n_events=2500):
    """the first n_events events of a fixture read and the signal they cover (keeps the variant files small)"""
    with h5mini.File(fn) as f:
        g = f["/Analyses/Basecall_1D_000"]
        rd_name, rd = list(f["/Raw/Reads"].items())[0]
        ev = f["/Analyses/Basecall_1D_000/BaseCalled_template/Events"][()][:n_events]
        sig = rd["Signal"][()][:int(ev["start"][-1] + ev["length"][-1]) + 37]
        return dict(version=bytes(g.attrs["version"]), events=ev,
                    fastq=f["/Analyses/Basecall_1D_000/BaseCalled_template/Fastq"][()], read_name=rd_name,
                    read_attrs={k: v for k, v in rd.attrs.items() if np.asarray(v).dtype.kind in "iuf"}, signal=sig)


def tree_single(fx, signal_ds, events=None, version=True):
    ana = {"BaseCalled_template": {"Events": events if events is not None else fx["events"],
                                   "Fastq": np.array(bytes(fx["fastq"]), dtype="S%d" % len(bytes(fx["fastq"])))}}
    if version:
        ana["__attrs__"] = {"version": fx["version"]}
    return {"Analyses": {"Basecall_1D_000": ana},
            "Raw": {"Reads": {fx["read_name"]: {"Signal": signal_ds, "__attrs__": fx["read_attrs"]}}}}


def main():
    os.makedirs(OUT, exist_ok=True)
    files = sorted(glob.glob(os.path.join(HERE, "fast5", "*.fast5")), key=os.path.getsize)
    fx, fx2 = load_fixture(files[0]), load_fixture(files[1])
    vbz = plugin_filter()
    chunk = 4096

    def put(name, tree):
        with open(os.path.join(OUT, name), "wb") as fp:
            fp.write(h5write.write_tree(tree))

    put("plain.fast5", tree_single(fx, (fx["signal"], dict(chunk=chunk, filt=("deflate", 1)))))
    for name, cd in (("vbz_v0.fast5", [0, 2, 1, 1]), ("vbz_v1.fast5", [1, 2, 1, 1]), ("vbz_nozstd.fast5", [0, 2, 1, 0])):
        put(name, tree_single(fx, (fx["signal"], dict(chunk=chunk, filt=(32020, cd, "vbz"), encode=lambda b, cd=cd: vbz(b, cd)))))
    # legacy table: float seconds, rescaled by the reader as start * 4000 - start_time (then int() truncation, :93)
    ev = fx["events"]
    st_time = int(fx["read_attrs"]["start_time"])
    ldt = np.dtype([(n, ("<f8" if n in ("start", "length") else ev.dtype[n])) for n in ev.dtype.names])
    lev = np.zeros(len(ev), dtype=ldt)
    for n in ev.dtype.names:
        lev[n] = ev[n]
    lev["start"] = (ev["start"].astype(np.float64) + st_time) / 4000.0
    lev["length"] = ev["length"].astype(np.float64) / 4000.0
    put("legacy.fast5", tree_single(fx, (fx["signal"], dict(chunk=chunk, filt=("deflate", 1))), events=lev, version=False))
    # multi-read container (two reads; one VBZ, one deflate)
    multi = {}
    for k, (f_, filt, enc) in enumerate(((fx, (32020, [0, 2, 1, 1], "vbz"), lambda b: vbz(b, [0, 2, 1, 1])), (fx2, ("deflate", 1), None))):
        kw = dict(chunk=chunk, filt=filt)
        if enc:
            kw["encode"] = enc
        multi["read_%08d-0000-4000-8000-%012d" % (k + 1, k + 1)] = {
            "Raw": {"Signal": (f_["signal"], kw), "__attrs__": f_["read_attrs"]},
            "Analyses": {"Basecall_1D_000": {"__attrs__": {"version": f_["version"]}, "BaseCalled_template": {
                "Events": f_["events"], "Fastq": np.array(bytes(f_["fastq"]), dtype="S%d" % len(bytes(f_["fastq"])))}}}}
    multi["__attrs__"] = {"file_version": b"2.0"}
    put("multi.fast5", multi)
    # decoder unit vectors from the plugin
    rng = np.random.default_rng(3)
    cases = {"real": fx["signal"][:5000], "random": rng.integers(-32768, 32768, 3000).astype(np.int16),
             "extremes": np.array([32767, -32768, -32768, 32767, 0, -1, 1, 32767, 32767, -32768], np.int16),
             "small": rng.integers(-3, 4, 1001).astype(np.int16), "one": np.array([-7], np.int16)}
    vec = {}
    for nm, a in cases.items():
        for tag, cd in (("v0z1", [0, 2, 1, 1]), ("v1z1", [1, 2, 1, 1]), ("v0z0", [0, 2, 1, 0]), ("nozz", [0, 2, 0, 1])):
            c = vbz(a.tobytes(), cd)
            assert vbz(c, cd, reverse=True) == a.tobytes()
            vec["%s_%s_raw" % (nm, tag)] = a
            vec["%s_%s_cd" % (nm, tag)] = np.array(cd, np.uint32)
            vec["%s_%s_comp" % (nm, tag)] = np.frombuffer(c, np.uint8)
    np.savez_compressed(os.path.join(OUT, "vbz_chunks.npz"), **vec)
    for fn in sorted(os.listdir(OUT)):
        print("%-20s %8d bytes" % (fn, os.path.getsize(os.path.join(OUT, fn))))


if __name__ == "__main__":
    main()
