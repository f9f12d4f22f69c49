"""ctypes binding of libnrv.so (include/nrv.h) and the batch packing around it.

This is the thin host layer of the hot path: it owns no arithmetic.  If the CUDA library is
missing or no B200 is visible it raises -- there is no CPU fallback and nothing here imports
``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from .weights import ModelWeights

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libnrv.so")

NRV_READ_OK, NRV_READ_TOO_SHORT, NRV_READ_SCALE_ZERO, NRV_READ_BAD_EVENTS = 0, 1, 2, 3


class NrvError(RuntimeError):
    pass


class _LstmDir(C.Structure):
    _fields_ = [("kernel", C.c_void_p), ("recurrent", C.c_void_p), ("bias", C.c_void_p)]


class _ModelWeights(C.Structure):
    _fields_ = [("window", C.c_int32), ("n_class", C.c_int32),
                ("conv1_k", C.c_void_p), ("conv1_b", C.c_void_p), ("bn1", C.c_void_p),
                ("conv2_k", C.c_void_p), ("conv2_b", C.c_void_p), ("bn2", C.c_void_p),
                ("sig_dense_k", C.c_void_p), ("sig_dense_b", C.c_void_p),
                ("lstm", (_LstmDir * 2) * 4),
                ("bn_rnn", C.c_void_p * 3),
                ("dense1_k", C.c_void_p), ("dense1_b", C.c_void_p),
                ("dense2_k", C.c_void_p), ("dense2_b", C.c_void_p),
                ("main_k", C.c_void_p), ("main_b", C.c_void_p),
                ("feat_k", C.c_void_p), ("feat_b", C.c_void_p),
                ("final_k", C.c_void_p), ("final_b", C.c_void_p)]


class _Batch(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("signal", C.c_void_p), ("sig_off", C.c_void_p),
                ("starts", C.c_void_p), ("base_off", C.c_void_p), ("bases", C.c_void_p),
                ("ev_mean", C.c_void_p), ("ev_std", C.c_void_p), ("last_dur", C.c_void_p), ("qual", C.c_void_p)]


class _Result(C.Structure):
    _fields_ = [("revised", C.c_void_p), ("revised_cap", C.c_int64), ("out_off", C.c_void_p),
                ("status", C.c_void_p), ("y1", C.c_void_p), ("y2", C.c_void_p),
                ("p1", C.c_void_p), ("p2", C.c_void_p), ("revised_qual", C.c_void_p)]


EXPORTS = ("nrv_create", "nrv_destroy", "nrv_last_error", "nrv_version", "nrv_launch_count",
           "nrv_stage_count", "nrv_stage_name", "nrv_set_stage_timing", "nrv_get_stage_ms", "nrv_get_stage_launches", "nrv_stream", "nrv_synchronize", "nrv_segment",
           "nrv_predict_windows", "nrv_decode", "nrv_revise_batch", "nrv_revise_batch_device", "nrv_submit_batch",
           "nrv_wait_batch", "nrv_debug_gemm",
           "nrv_ingest_fast5", "nrv_ingest_view", "nrv_ingest_read_names", "nrv_ingest_free")

_lib = None


def load_library(path: Optional[str] = None):
    """dlopen libnrv.so and declare the prototypes of every symbol in include/nrv.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise NrvError("libnrv.so not built (%s): run `python -m nanoreviser_b200.build` "
                       "or __graft_entry__.build(); there is no CPU fallback" % p)
    lib = C.CDLL(p)
    vp = C.c_void_p
    lib.nrv_create.argtypes = [C.c_int, C.POINTER(_ModelWeights), C.POINTER(_ModelWeights), C.POINTER(vp)]
    lib.nrv_create.restype = C.c_int
    lib.nrv_destroy.argtypes = [vp]
    lib.nrv_destroy.restype = None
    lib.nrv_last_error.argtypes = [vp]
    lib.nrv_last_error.restype = C.c_char_p
    lib.nrv_version.argtypes = []
    lib.nrv_version.restype = C.c_char_p
    lib.nrv_launch_count.argtypes = [vp]
    lib.nrv_launch_count.restype = C.c_int64
    lib.nrv_set_stage_timing.argtypes = [vp, C.c_int]
    lib.nrv_set_stage_timing.restype = C.c_int
    lib.nrv_stage_count.argtypes = []
    lib.nrv_stage_count.restype = C.c_int
    lib.nrv_stage_name.argtypes = [C.c_int]
    lib.nrv_stage_name.restype = C.c_char_p
    lib.nrv_get_stage_ms.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
    lib.nrv_get_stage_ms.restype = C.c_int
    lib.nrv_get_stage_launches.argtypes = [vp, C.POINTER(C.c_int64), C.c_int]
    lib.nrv_get_stage_launches.restype = C.c_int
    lib.nrv_stream.argtypes = [vp]
    lib.nrv_stream.restype = vp
    lib.nrv_synchronize.argtypes = [vp]
    lib.nrv_synchronize.restype = C.c_int
    lib.nrv_segment.argtypes = [vp, C.POINTER(_Batch), vp, vp, vp, vp, vp, vp, vp]
    lib.nrv_segment.restype = C.c_int
    lib.nrv_predict_windows.argtypes = [vp, C.c_int64, vp, vp, vp, vp]
    lib.nrv_predict_windows.restype = C.c_int
    lib.nrv_decode.argtypes = [vp, C.c_int64, vp, vp, vp, vp, vp, vp, C.c_int64, vp]
    lib.nrv_decode.restype = C.c_int
    lib.nrv_revise_batch.argtypes = [vp, C.POINTER(_Batch), C.POINTER(_Result)]
    lib.nrv_revise_batch.restype = C.c_int
    lib.nrv_revise_batch_device.argtypes = [vp, C.POINTER(_Batch), C.POINTER(_Result)]
    lib.nrv_revise_batch_device.restype = C.c_int
    lib.nrv_submit_batch.argtypes = [vp, C.POINTER(_Batch), C.POINTER(_Result), C.POINTER(C.c_int64)]
    lib.nrv_submit_batch.restype = C.c_int
    lib.nrv_wait_batch.argtypes = [vp, C.c_int64]
    lib.nrv_wait_batch.restype = C.c_int
    lib.nrv_debug_gemm.argtypes = [vp, C.c_int64, C.c_int, C.c_int, vp, vp, vp, vp]
    lib.nrv_debug_gemm.restype = C.c_int
    lib.nrv_ingest_fast5.argtypes = [C.POINTER(C.c_char_p), C.c_int64, C.c_char_p, C.c_char_p, C.c_int, C.POINTER(vp)]
    lib.nrv_ingest_fast5.restype = C.c_int
    lib.nrv_ingest_view.argtypes = [vp, C.POINTER(_Batch), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.nrv_ingest_view.restype = C.c_int
    lib.nrv_ingest_read_names.argtypes = [vp, C.POINTER(C.POINTER(C.c_char_p))]
    lib.nrv_ingest_read_names.restype = C.c_int
    lib.nrv_ingest_free.argtypes = [vp]
    lib.nrv_ingest_free.restype = None
    if path is None:
        _lib = lib
    return lib


INGEST_OK, INGEST_OPEN_FAILED, INGEST_NO_EVENTS, INGEST_TOO_SHORT, INGEST_NO_SIGNAL, INGEST_SIGNAL_SHORT, INGEST_CORRUPT, \
    INGEST_UNSUPPORTED = range(8)


def ingest_fast5(paths: Sequence[str], basecall_group: str = "Basecall_1D_000", basecall_subgroup: str = "BaseCalled_template",
                 threads: int = 0, with_names: bool = False):
    """Native multi-threaded ``get_read_data`` over a list of fast5 files (include/nrv.h: nrv_ingest_fast5): single-read files
    (deflate or VBZ signal, current or legacy event tables) and multi-read containers (one read per member).

    -> ``(batch, file_status int32[n_files], read_file int64[n_reads], a0 int64[n_reads])``: ``batch`` holds the reads that
    decoded, in file order; the arrays are copies owned by numpy.  ``with_names=True`` appends ``names``: the member name of
    every read that came out of a multi-read container, ``""`` for single-read files.  Host-only: works without a GPU."""
    lib = load_library()
    n = len(paths)
    arr = (C.c_char_p * max(n, 1))(*[os.fsencode(p) for p in paths])
    h = C.c_void_p()
    rc = lib.nrv_ingest_fast5(arr, n, basecall_group.encode(), basecall_subgroup.encode(), int(threads), C.byref(h))
    if rc != 0:
        raise NrvError("nrv_ingest_fast5 failed (%d)" % rc)
    try:
        cb = _Batch()
        fs, rf, a0 = C.c_void_p(), C.c_void_p(), C.c_void_p()
        rc = lib.nrv_ingest_view(h, C.byref(cb), C.byref(fs), C.byref(rf), C.byref(a0))
        if rc != 0:
            raise NrvError("nrv_ingest_view failed (%d)" % rc)
        R = int(cb.n_reads)

        def view(ptr, count, dt):
            if count == 0 or not ptr:
                return np.zeros(0, dt)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dt))), shape=(count,)).copy()
        sig_off = view(cb.sig_off, R + 1, np.int64)
        base_off = view(cb.base_off, R + 1, np.int64)
        nb, ns = int(base_off[-1]), int(sig_off[-1])
        batch = Batch(view(cb.signal, ns, np.int16), sig_off, view(cb.starts, nb, np.int32), base_off,
                      view(cb.bases, nb, np.uint8), view(cb.ev_mean, nb, np.float32), view(cb.ev_std, nb, np.float32),
                      view(cb.last_dur, R, np.int32), view(cb.qual, nb, np.uint8) if cb.qual else None)
        out = (batch, view(fs.value, n, np.int32), view(rf.value, R, np.int64), view(a0.value, R, np.int64))
        if with_names:
            np_ = C.POINTER(C.c_char_p)()
            if lib.nrv_ingest_read_names(h, C.byref(np_)) != 0:
                raise NrvError("nrv_ingest_read_names failed")
            out = out + ([np_[i].decode() for i in range(R)],)
        return out
    finally:
        lib.nrv_ingest_free(h)


def _ptr(a) -> Optional[int]:
    if a is None:
        return None
    return a.ctypes.data


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _pack_weights(m: ModelWeights, keep: list) -> _ModelWeights:
    w = _ModelWeights()
    w.window, w.n_class = int(m.window), int(m.n_class)

    def put(name, arr):
        a = _f32(arr)
        keep.append(a)
        setattr(w, name, a.ctypes.data)

    for name in ("conv1_k", "conv1_b", "bn1", "conv2_k", "conv2_b", "bn2", "sig_dense_k", "sig_dense_b",
                 "dense1_k", "dense1_b", "dense2_k", "dense2_b", "main_k", "main_b", "feat_k", "feat_b",
                 "final_k", "final_b"):
        put(name, getattr(m, name))
    for l in range(4):
        for d in range(2):
            src = m.lstm[l][d]
            for fld in ("kernel", "recurrent", "bias"):
                a = _f32(getattr(src, fld))
                keep.append(a)
                setattr(w.lstm[l][d], fld, a.ctypes.data)
    for i in range(3):
        a = _f32(m.bn_rnn[i])
        keep.append(a)
        w.bn_rnn[i] = a.ctypes.data
    return w


@dataclass
class Batch:
    """Ragged CSR batch in the layout of include/nrv.h."""
    signal: np.ndarray      # int16 [sum S]   (signal[a0:] of every read)
    sig_off: np.ndarray     # int64 [R+1]
    starts: np.ndarray      # int32 [sum N]
    base_off: np.ndarray    # int64 [R+1]
    bases: np.ndarray       # uint8 [sum N]
    ev_mean: np.ndarray     # float32 [sum N]
    ev_std: np.ndarray      # float32 [sum N]
    last_dur: np.ndarray    # int32 [R]
    qual: Optional[np.ndarray] = None   # uint8 [sum N], optional: basecaller Phred scores (Fastq quality char - 33)

    @property
    def n_reads(self) -> int:
        return int(self.last_dur.shape[0])

    @property
    def n_bases(self) -> int:
        return int(self.base_off[-1])

    def n_windows(self, window: int) -> int:
        n = np.diff(self.base_off) - window
        return int(np.maximum(n, 0).sum())

    def win_off(self, window: int) -> np.ndarray:
        out = np.zeros(self.n_reads + 1, dtype=np.int64)
        np.cumsum(np.maximum(np.diff(self.base_off) - window, 0), out=out[1:])
        return out

    def h2d_bytes(self) -> int:
        return int(self.signal.nbytes + self.sig_off.nbytes + self.starts.nbytes + self.base_off.nbytes +
                   self.bases.nbytes + self.ev_mean.nbytes + self.ev_std.nbytes + self.last_dur.nbytes)


def pack_batch(reads: Sequence) -> Batch:
    """reads: objects with .signal (whole raw signal) .a0 .starts .bases .ev_mean .ev_std .last_dur
    (``fast5.ReadArrays``) -> Batch.  Applies ``signal = or_raw_signal[a0:]`` (NanoReviser.py:120)."""
    R = len(reads)
    sig_off = np.zeros(R + 1, dtype=np.int64)
    base_off = np.zeros(R + 1, dtype=np.int64)
    for i, r in enumerate(reads):
        sig_off[i + 1] = sig_off[i] + (len(r.signal) - int(r.a0))
        base_off[i + 1] = base_off[i] + len(r.starts)
    signal = np.empty(int(sig_off[-1]), dtype=np.int16)
    starts = np.empty(int(base_off[-1]), dtype=np.int32)
    bases = np.empty(int(base_off[-1]), dtype=np.uint8)
    evm = np.empty(int(base_off[-1]), dtype=np.float32)
    evs = np.empty(int(base_off[-1]), dtype=np.float32)
    last = np.empty(R, dtype=np.int32)
    for i, r in enumerate(reads):
        signal[sig_off[i]:sig_off[i + 1]] = r.signal[int(r.a0):]
        s, e = base_off[i], base_off[i + 1]
        starts[s:e] = r.starts
        bases[s:e] = r.bases
        evm[s:e] = r.ev_mean
        evs[s:e] = r.ev_std
        last[i] = int(r.last_dur)
    qual = None
    if R and all(getattr(r, "qual", None) is not None for r in reads):
        qual = np.concatenate([np.asarray(r.qual, dtype=np.uint8) for r in reads])
    return Batch(signal, sig_off, starts, base_off, bases, evm, evs, last, qual)


def slice_batch(b: Batch, r0: int, r1: int) -> Batch:
    """Reads [r0, r1) of a batch as VIEWS of its arrays (no copy; only the two offset arrays are rebased)."""
    s0, s1 = int(b.sig_off[r0]), int(b.sig_off[r1])
    b0, b1 = int(b.base_off[r0]), int(b.base_off[r1])
    return Batch(b.signal[s0:s1], b.sig_off[r0:r1 + 1] - s0, b.starts[b0:b1], b.base_off[r0:r1 + 1] - b0, b.bases[b0:b1],
                 b.ev_mean[b0:b1], b.ev_std[b0:b1], b.last_dur[r0:r1], None if b.qual is None else b.qual[b0:b1])


def split_batch(b: Batch, idx: Sequence[int]) -> Batch:
    """Sub-batch with the given reads (in the given order)."""
    idx = list(idx)
    R = len(idx)
    sig_off = np.zeros(R + 1, dtype=np.int64)
    base_off = np.zeros(R + 1, dtype=np.int64)
    sig, st, ba, em, es = [], [], [], [], []
    for k, i in enumerate(idx):
        s0, s1 = int(b.sig_off[i]), int(b.sig_off[i + 1])
        b0, b1 = int(b.base_off[i]), int(b.base_off[i + 1])
        sig_off[k + 1] = sig_off[k] + (s1 - s0)
        base_off[k + 1] = base_off[k] + (b1 - b0)
        sig.append(b.signal[s0:s1]); st.append(b.starts[b0:b1]); ba.append(b.bases[b0:b1])
        em.append(b.ev_mean[b0:b1]); es.append(b.ev_std[b0:b1])
    cat = lambda parts, dt: (np.concatenate(parts).astype(dt, copy=False) if parts else np.zeros(0, dt))
    qual = None
    if getattr(b, "qual", None) is not None:
        qual = cat([b.qual[int(b.base_off[i]):int(b.base_off[i + 1])] for i in idx], np.uint8)
    return Batch(cat(sig, np.int16), sig_off, cat(st, np.int32), base_off, cat(ba, np.uint8), cat(em, np.float32),
                 cat(es, np.float32), np.asarray(b.last_dur)[idx].astype(np.int32), qual)


@dataclass
class ReviseResult:
    revised: np.ndarray     # uint8, concatenated
    out_off: np.ndarray     # int64 [R+1]
    status: np.ndarray      # int32 [R]
    y1: Optional[np.ndarray] = None
    y2: Optional[np.ndarray] = None
    p1: Optional[np.ndarray] = None
    p2: Optional[np.ndarray] = None
    revised_qual: Optional[np.ndarray] = None   # uint8 Phred+33 characters parallel to ``revised`` (want_qual)

    def sequence(self, i: int) -> str:
        return self.revised[self.out_off[i]:self.out_off[i + 1]].tobytes().decode("ascii")

    def quality(self, i: int) -> str:
        """Fastq quality string of read ``i`` (definition D6', include/nrv.h nrv_result.revised_qual)."""
        if self.revised_qual is None:
            raise ValueError("revise_batch was called without want_qual")
        return self.revised_qual[self.out_off[i]:self.out_off[i + 1]].tobytes().decode("ascii")

    def sequences(self) -> List[str]:
        return [self.sequence(i) for i in range(len(self.status))]


class Reviser:
    """One handle per GPU (include/nrv.h).  Not thread-safe; re-entrant across handles."""

    def __init__(self, model1: ModelWeights, model2: ModelWeights, device: int = 0, lib_path: Optional[str] = None):
        self._lib = load_library(lib_path)
        self._keep: list = []
        w1 = _pack_weights(model1, self._keep)
        w2 = _pack_weights(model2, self._keep)
        h = C.c_void_p()
        rc = self._lib.nrv_create(int(device), C.byref(w1), C.byref(w2), C.byref(h))
        if rc != 0:
            raise NrvError("nrv_create failed (%d): %s" % (rc, self._lib.nrv_last_error(None).decode()))
        self._h = h
        self.window = int(model1.window)
        self.device = int(device)

    # -- plumbing ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.nrv_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise NrvError("%s failed (%d): %s" % (what, rc, self._lib.nrv_last_error(self._h).decode()))

    @property
    def launch_count(self) -> int:
        return int(self._lib.nrv_launch_count(self._h))

    @property
    def stream(self) -> int:
        return int(self._lib.nrv_stream(self._h) or 0)

    def synchronize(self):
        self._check(self._lib.nrv_synchronize(self._h), "nrv_synchronize")

    def set_stage_timing(self, enable: bool):
        self._check(self._lib.nrv_set_stage_timing(self._h, int(enable)), "nrv_set_stage_timing")

    def stage_ms(self) -> dict:
        n = self._lib.nrv_stage_count()
        buf = (C.c_float * n)()
        self._check(self._lib.nrv_get_stage_ms(self._h, buf, n), "nrv_get_stage_ms")
        return {self._lib.nrv_stage_name(i).decode(): float(buf[i]) for i in range(n)}

    def stage_launches(self) -> dict:
        n = self._lib.nrv_stage_count()
        buf = (C.c_int64 * n)()
        self._check(self._lib.nrv_get_stage_launches(self._h, buf, n), "nrv_get_stage_launches")
        return {self._lib.nrv_stage_name(i).decode(): int(buf[i]) for i in range(n)}

    @staticmethod
    def _cbatch(b: Batch, keep: list) -> _Batch:
        cb = _Batch()
        cb.n_reads = b.n_reads
        arrs = dict(signal=np.ascontiguousarray(b.signal, dtype=np.int16),
                    sig_off=np.ascontiguousarray(b.sig_off, dtype=np.int64),
                    starts=np.ascontiguousarray(b.starts, dtype=np.int32),
                    base_off=np.ascontiguousarray(b.base_off, dtype=np.int64),
                    bases=np.ascontiguousarray(b.bases, dtype=np.uint8),
                    ev_mean=np.ascontiguousarray(b.ev_mean, dtype=np.float32),
                    ev_std=np.ascontiguousarray(b.ev_std, dtype=np.float32),
                    last_dur=np.ascontiguousarray(b.last_dur, dtype=np.int32))
        for k, a in arrs.items():
            keep.append(a)
            setattr(cb, k, a.ctypes.data)
        if getattr(b, "qual", None) is not None:
            q = np.ascontiguousarray(b.qual, dtype=np.uint8)
            if q.shape[0] != arrs["bases"].shape[0]:
                raise ValueError("Batch.qual must hold one Phred score per base")
            keep.append(q)
            cb.qual = q.ctypes.data
        return cb

    # -- stage-level entry points -------------------------------------------------------------
    def segment(self, b: Batch, want_windows: bool = False):
        """nrv_segment: (shift[R], scale[R], seg_mean[N], seg_std[N], x[N,6], sig_win[N,50] | None, status[R])."""
        keep: list = []
        cb = self._cbatch(b, keep)
        R, N = b.n_reads, b.n_bases
        shift = np.empty(R, np.float64); scale = np.empty(R, np.float64)
        mean = np.empty(N, np.float64); std = np.empty(N, np.float64)
        x = np.empty((N, 6), np.float32)
        win = np.empty((N, 50), np.float32) if want_windows else None
        status = np.empty(R, np.int32)
        rc = self._lib.nrv_segment(self._h, C.byref(cb), _ptr(shift), _ptr(scale), _ptr(mean), _ptr(std), _ptr(x),
                                   _ptr(win), _ptr(status))
        self._check(rc, "nrv_segment")
        return shift, scale, mean, std, x, win, status

    def predict_windows(self, S: np.ndarray, X: np.ndarray):
        """nrv_predict_windows: S [n,W,50], X [n,W,6] -> (P1 [n,6], P2 [n,5])."""
        S = np.ascontiguousarray(S, dtype=np.float32)
        X = np.ascontiguousarray(X, dtype=np.float32)
        if S.ndim == 4 and S.shape[-1] == 1:
            S = S[..., 0]
        n = S.shape[0]
        if S.shape != (n, self.window, 50) or X.shape != (n, self.window, 6):
            raise NrvError("expected S [n,%d,50] and X [n,%d,6]" % (self.window, self.window))
        p1 = np.empty((n, 6), np.float32)
        p2 = np.empty((n, 5), np.float32)
        self._check(self._lib.nrv_predict_windows(self._h, n, _ptr(S), _ptr(X), _ptr(p1), _ptr(p2)),
                    "nrv_predict_windows")
        return p1, p2

    def decode(self, base_off, bases, y1, y2, status=None):
        base_off = np.ascontiguousarray(base_off, dtype=np.int64)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        y1 = np.ascontiguousarray(y1, dtype=np.uint8)
        y2 = np.ascontiguousarray(y2, dtype=np.uint8)
        R = len(base_off) - 1
        st = None if status is None else np.ascontiguousarray(status, dtype=np.int32)
        cap = 2 * int(base_off[-1]) + R + 16
        revised = np.empty(cap, np.uint8)
        out_off = np.empty(R + 1, np.int64)
        rc = self._lib.nrv_decode(self._h, R, _ptr(base_off), _ptr(bases), _ptr(y1), _ptr(y2), _ptr(st),
                                  _ptr(revised), cap, _ptr(out_off))
        self._check(rc, "nrv_decode")
        return revised[:out_off[-1]], out_off

    def debug_gemm(self, A: np.ndarray, Bt: np.ndarray, bias: Optional[np.ndarray] = None) -> np.ndarray:
        """nrv_debug_gemm: A [M,K] . Bt[N,K]^T (+bias) through the tcgen05 split-fp16 projection GEMM."""
        A = np.ascontiguousarray(A, dtype=np.float32)
        Bt = np.ascontiguousarray(Bt, dtype=np.float32)
        M, K = A.shape
        N = Bt.shape[0]
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
        Cm = np.empty((M, N), np.float32)
        self._check(self._lib.nrv_debug_gemm(self._h, M, N, K, _ptr(A), _ptr(Bt), _ptr(b), _ptr(Cm)), "nrv_debug_gemm")
        return Cm

    # -- the whole path -----------------------------------------------------------------------
    def revise_batch(self, b: Batch, want_labels: bool = False, want_probs: bool = False,
                     out: Optional[ReviseResult] = None, want_qual: bool = False) -> ReviseResult:
        """nrv_revise_batch: submit + wait."""
        return self.wait(self.submit(b, want_labels, want_probs, out, want_qual))

    def submit(self, b: Batch, want_labels: bool = False, want_probs: bool = False,
               out: Optional[ReviseResult] = None, want_qual: bool = False):
        """nrv_submit_batch: enqueue one host batch (H2D on the copy stream, kernels on the main stream) and return a pending
        object for :meth:`wait`.  Up to two batches may be in flight: submit batch i+1, then wait for batch i -- the copies of
        one run under the kernels of the other.  The arrays of ``b`` must stay alive (and unmodified) until ``wait``."""
        keep: list = [b]
        cb = self._cbatch(b, keep)
        R, N = b.n_reads, b.n_bases
        nw = b.n_windows(self.window)
        cap = 2 * N + R + 16
        if out is None:
            out = ReviseResult(np.empty(cap, np.uint8), np.empty(R + 1, np.int64), np.empty(R, np.int32))
            if want_labels:
                out.y1 = np.empty(nw, np.uint8); out.y2 = np.empty(nw, np.uint8)
            if want_probs:
                out.p1 = np.empty((nw, 6), np.float32); out.p2 = np.empty((nw, 5), np.float32)
            if want_qual:
                out.revised_qual = np.empty(cap, np.uint8)
        cr = _Result()
        cr.revised = _ptr(out.revised); cr.revised_cap = int(out.revised.shape[0])
        cr.out_off = _ptr(out.out_off); cr.status = _ptr(out.status)
        cr.y1 = _ptr(out.y1); cr.y2 = _ptr(out.y2); cr.p1 = _ptr(out.p1); cr.p2 = _ptr(out.p2)
        cr.revised_qual = _ptr(out.revised_qual)
        ticket = C.c_int64(0)
        self._check(self._lib.nrv_submit_batch(self._h, C.byref(cb), C.byref(cr), C.byref(ticket)), "nrv_submit_batch")
        return (int(ticket.value), out, keep)

    def wait(self, pending) -> ReviseResult:
        """nrv_wait_batch: block until the batch's results are in its ReviseResult (D2H on the copy stream)."""
        ticket, out, _keep = pending
        self._check(self._lib.nrv_wait_batch(self._h, ticket), "nrv_wait_batch")
        return out

    def revise_batch_device(self, n_reads: int, sig_off: np.ndarray, base_off: np.ndarray, dptr: dict, dres: dict,
                            revised_cap: int):
        """nrv_revise_batch_device: ``dptr`` / ``dres`` map field name -> raw device pointer (int).
        Offsets stay host arrays.  Enqueues on ``self.stream`` and returns without synchronising."""
        cb = _Batch()
        cb.n_reads = int(n_reads)
        sig_off = np.ascontiguousarray(sig_off, dtype=np.int64)
        base_off = np.ascontiguousarray(base_off, dtype=np.int64)
        cb.sig_off = sig_off.ctypes.data
        cb.base_off = base_off.ctypes.data
        for k in ("signal", "starts", "bases", "ev_mean", "ev_std", "last_dur"):
            setattr(cb, k, int(dptr[k]))
        cr = _Result()
        cr.revised = int(dres["revised"]); cr.revised_cap = int(revised_cap)
        cr.out_off = int(dres["out_off"]); cr.status = int(dres["status"])
        cb.qual = int(dptr["qual"]) if dptr.get("qual") else None
        for k in ("y1", "y2", "p1", "p2", "revised_qual"):
            v = dres.get(k)
            setattr(cr, k, int(v) if v else None)
        self._check(self._lib.nrv_revise_batch_device(self._h, C.byref(cb), C.byref(cr)), "nrv_revise_batch_device")
