"""Host-side fast5 ingest (row A1 of SURVEY.md section 8): event table -> per-base arrays.

Mirrors ``get_read_data`` / ``extract_fastq`` of the reference
(``nanorevutils/nanorev_fast5_handeler.py:39-150`` and ``:152-171``): same names,
argument meaning, return tuple and error behaviour, but the per-event Python loop
(``:84-118``) is a vectorised flag -> scan -> scatter over the packed ``Events`` table
and the HDF5 access goes through :mod:`nanoreviser_b200.h5mini`.

The output of this stage is the ragged H2D payload of the CUDA path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import h5mini


@dataclass
class ReadArrays:
    """Per-read arrays in the layout the C-ABI consumes (see include/nrv.h)."""
    name: str
    a0: int                 # abs_event_start (start of the first base in the raw signal)
    starts: np.ndarray      # int64[N], relative to a0
    length: np.ndarray      # float64[N]  (as the reference returns it)
    bases: np.ndarray       # uint8[N] ASCII
    signal: np.ndarray      # int16[S_total] (whole raw signal; the hot path uses signal[a0:])
    ev_mean: np.ndarray     # float32[N]
    ev_std: np.ndarray      # float32[N]
    qual: "np.ndarray | None" = None   # uint8[N] optional: basecaller Phred scores (see basecall_phred)

    @property
    def n_bases(self) -> int:
        return int(self.starts.shape[0])

    @property
    def last_dur(self) -> int:
        return int(self.length[-1])


def _version_le_zero(v) -> bool:
    """``LooseVersion(v) <= LooseVersion('0.0')`` for dotted numeric versions (fast5_handeler.py:65-68)."""
    if isinstance(v, bytes):
        v = v.decode()
    parts = []
    for p in str(v).replace("-", ".").split("."):
        try:
            parts.append(int(p))
        except ValueError:
            parts.append(1)    # any alphabetic tag sorts above '0.0'
    while len(parts) < 2:
        parts.append(0)
    return parts <= [0, 0]


def collapse_events(ev_start, ev_mean, ev_stdv, ev_state, ev_move):
    """Vectorised restatement of the reversed event loop (fast5_handeler.py:84-118).

    move == 0 -> skipped; move == 2 -> two bases ``(start, state[1])`` then ``(start + 2, state[2])``;
    every other move value -> one base ``(start, state[2])``.  Both bases of a move-2 event carry
    the event's mean / stdv.
    """
    ev_move = np.asarray(ev_move)
    n_out = np.where(ev_move == 0, 0, np.where(ev_move == 2, 2, 1)).astype(np.int64)
    off = np.zeros(len(n_out) + 1, dtype=np.int64)
    np.cumsum(n_out, out=off[1:])
    total = int(off[-1])
    src = np.repeat(np.arange(len(n_out), dtype=np.int64), n_out)    # event index per base
    rank = np.arange(total, dtype=np.int64) - off[src]               # 0 or 1 within the event
    two = ev_move[src] == 2
    st = np.asarray(ev_start)[src]
    # int(start_l) truncates (legacy float starts), as fast5_handeler.py:93
    st = st.astype(np.int64) if st.dtype.kind != "f" else np.trunc(st).astype(np.int64)
    start = st + np.where(two & (rank == 1), 2, 0)
    ev_state = np.ascontiguousarray(ev_state)
    width = ev_state.dtype.itemsize                 # k-mer width of the basecaller's model (5 for Albacore r9.4; any width >= 3:
    if width < 3:                                   # the reference indexes characters 1 and 2 whatever the length, :98-109)
        raise RuntimeError("model_state is narrower than 3 characters")
    states = np.frombuffer(ev_state.tobytes(), dtype=np.uint8).reshape(-1, width)
    col = np.where(two & (rank == 0), 1, 2)
    bases = states[src, col]
    return start, bases, np.asarray(ev_mean)[src].astype(np.float32), np.asarray(ev_stdv)[src].astype(np.float32)


def list_members(fast5_fn):
    """Members of a multi-read container (root groups named read_<id>), in name order; [] for a single-read file."""
    with h5mini.File(fast5_fn, "r") as f:
        keys = f.keys()
    if "Raw" in keys or "Analyses" in keys:
        return []
    return sorted(k for k in keys if k.startswith("read_"))


def read_fast5_arrays(fast5_fn, basecall_group="Basecall_1D_000",
                      basecall_subgroup="BaseCalled_template", member=None) -> ReadArrays:
    """``member``: a read_<id> group of a multi-read container (an extension: the reference opens single-read files only,
    fast5_handeler.py:132-133); its layout is <member>/Analyses/... and <member>/Raw/Signal."""
    try:
        f = h5mini.File(fast5_fn, "r")
    except Exception:
        raise NotImplementedError("Error opening file. Likely a corrupted file.")
    top = "/" + member if member else ""
    try:
        grp = f[top + "/Analyses/" + basecall_group]
        ver = grp.attrs["version"] if "version" in grp.attrs else "0.0"
        called = f[top + "/Analyses/" + basecall_group + "/" + basecall_subgroup + "/Events"][()]
        ev_start = called["start"]
        if _version_le_zero(ver):
            raw = f[top + "/Raw"] if member else list(f["/Raw/Reads/"].values())[0]
            raw_attrs = dict(raw.attrs.items())
            ev_start = (ev_start * 4000 - raw_attrs["start_time"]).astype(ev_start.dtype)
    except Exception:
        f.close()
        raise RuntimeError("No events or corrupted events in file. Likely a segmentation error .")
    start, bases, ab_mean, ab_std = collapse_events(ev_start, called["mean"], called["stdv"],
                                                    called["model_state"], called["move"])
    if len(start) < 2:
        f.close()
        raise RuntimeError("Events is too short or there are too much zero moves.")
    length = np.empty(len(start), dtype=np.float64)
    length[:-1] = np.diff(start)
    length[-1] = 3.0 if start[-1] - start[-2] < 5 else 5.0
    try:
        if member:
            signal = f[top + "/Raw/Signal"][()]
        else:
            read_name = list(f["/Raw/Reads/"].items())[0][0]
            signal = f["/Raw/Reads/" + str(read_name) + "/Signal"][()]
    except Exception:
        f.close()
        raise RuntimeError("No signal stored in the file")
    f.close()
    if len(signal) < int(start[-1] + length[-1]):
        raise RuntimeError("Signal is shorter than the Events")
    a0 = int(start[0])
    import os
    return ReadArrays(name=member if member else os.path.basename(str(fast5_fn)), a0=a0, starts=start - a0, length=length,
                      bases=bases, signal=np.ascontiguousarray(signal, dtype=np.int16),
                      ev_mean=ab_mean, ev_std=ab_std)


def get_read_data(fast5_fn, basecall_group, basecall_subgroup):
    """Same return tuple as the reference (fast5_handeler.py:145-150):
    ``(abs_event_start, start, length, bases, signal, ab_mean, ab_std)``."""
    r = read_fast5_arrays(fast5_fn, basecall_group, basecall_subgroup)
    bases = [chr(c) for c in r.bases]
    return r.a0, r.starts, r.length, bases, r.signal, list(r.ev_mean), list(r.ev_std)


def basecall_phred(fast5_fn, bases, basecall_group="Basecall_1D_000", basecall_subgroup="BaseCalled_template"):
    """Phred scores (uint8, quality character - 33) of the event-collapsed ``bases`` of a read, taken from the
    basecaller's Fastq dataset: the collapsed call is the Fastq sequence without its first and last two bases
    (``bases == Fastq_seq[2:-2]``, the Albacore known answer of SURVEY.md section 8(c)).  Returns None when the
    dataset is missing or does not line up -- the revision path then uses its constant pass-through quality."""
    try:
        f = h5mini.File(fast5_fn, "r")
        fastq = f["/Analyses/" + basecall_group + "/" + basecall_subgroup + "/Fastq"][()]
        f.close()
        lines = bytes(fastq).decode("utf8").split("\n")
        seq, qul = lines[1], lines[3]
    except Exception:
        return None
    b = bytes(bytearray(np.asarray(bases, dtype=np.uint8))).decode("ascii") if not isinstance(bases, str) else bases
    if len(seq) != len(qul) or len(seq) != len(b) + 4 or seq[2:-2] != b:
        return None
    q = np.frombuffer(qul[2:-2].encode("ascii"), dtype=np.uint8).astype(np.int16) - 33
    return np.clip(q, 0, 93).astype(np.uint8)


def extract_fastq(fast5_fn, out_fasta_fn=None, basecall_group="Basecall_1D_000",
                  basecall_subgroup="BaseCalled_template"):
    """Original Albacore call trimmed by 7 on both sides (fast5_handeler.py:152-171)."""
    try:
        f = h5mini.File(fast5_fn, "r")
    except Exception:
        raise NotImplementedError("Error opening file. Likely a corrupted file.")
    try:
        fastq = f["/Analyses/" + basecall_group + "/" + basecall_subgroup + "/Fastq"][()]
        f.close()
        lines = bytes(fastq).decode("utf8").split("\n")
        bases, qul = lines[1], lines[3]
        assert len(bases) >= 14
        assert len(bases) == len(qul)
        return bases[7:-7], qul[7:-7]
    except Exception:
        raise NotImplementedError("Error opening file. Likely a corrupted file.")
