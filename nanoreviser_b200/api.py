"""Host-side mirror of the reference's Python interface for the revision path.

Same names, argument meaning, return shapes and error behaviour as the functions that
``provide_fasta`` (NanoReviser.py:105-183) composes, so that the parity tests read like the
reference's own call sites -- but every computation runs in libnrv.so on the GPU:

=========================  =====================================================  ===================
this module                reference                                              C-ABI entry point
=========================  =====================================================  ===================
``get_read_data``          nanorev_fast5_handeler.py:39-150                       (host ingest)
``signal_segmentation``    preprocessing.py:85-170                                ``nrv_segment``
``get_model1/2().predict`` lstmmodel.py:32-133 + Keras ``Model.predict``          ``nrv_predict_windows``
``get_base_1``             output_handeler.py:104-122                             ``nrv_decode``
``prep_read_fasta/fastq``  output_handeler.py:26-62                               (host writer)
``revise_reads``           body of provide_fasta, NN path (composition D1-D8)     ``nrv_revise_batch``
=========================  =====================================================  ===================
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np

from . import engine, weights
from .fast5 import ReadArrays, extract_fastq, get_read_data, read_fast5_arrays  # noqa: F401 (re-exported)

_default: Optional[engine.Reviser] = None


def init(species: str = "human", model_root: str = "./model", device: int = 0) -> engine.Reviser:
    """Load ``./model/<S>/<S>_win13_50ep_model{1,2}.h5`` (NanoReviser.py:192-193) onto one GPU."""
    global _default
    m1, m2 = weights.load_species(species, model_root)
    _default = engine.Reviser(m1, m2, device=device)
    return _default


def set_default(r: Optional[engine.Reviser]):
    global _default
    _default = r


def _rev(r: Optional[engine.Reviser]) -> engine.Reviser:
    r = r or _default
    if r is None:
        raise engine.NrvError("no CUDA handle: call nanoreviser_b200.api.init(species) first")
    return r


# ---- preprocessing.py ---------------------------------------------------------------------------
def get_base_color(base):
    return {'A': 250, 'G': 180, 'T': 100, 'C': 30}.get(base, 0)


def get_base_label(base):
    return {'A': 5, 'G': 4, 'T': 3, 'C': 2, '-': 1, 'D': 0}.get(base, 0)


def signal_segmentation(raw_signal, starts, last_dur, query_len=50, reviser: Optional[engine.Reviser] = None):
    """-> (signal_list [N,50], signal_mean list, signal_std list, shift, scale)  (preprocessing.py:170).

    Windows come back as the fp32 values the model consumes; mean/std/shift/scale are float64 like the
    reference's.  Raises ``RuntimeError('Signal segmentation Error')`` like preprocessing.py:168-169.
    """
    if int(query_len) != 50:
        raise RuntimeError('Signal segmentation Error')          # the model input is fixed at 50 (lstmmodel.py:28)
    r = _rev(reviser)
    sig = np.ascontiguousarray(raw_signal, dtype=np.int16)
    st = np.ascontiguousarray(starts, dtype=np.int32)
    n = st.shape[0]
    b = engine.Batch(signal=sig, sig_off=np.array([0, sig.shape[0]], np.int64), starts=st,
                     base_off=np.array([0, n], np.int64), bases=np.full(n, ord('A'), np.uint8),
                     ev_mean=np.zeros(n, np.float32), ev_std=np.zeros(n, np.float32),
                     last_dur=np.array([int(last_dur)], np.int32))
    shift, scale, mean, std, _x, win, status = r.segment(b, want_windows=True)
    if status[0] == engine.NRV_READ_BAD_EVENTS:
        raise RuntimeError('Signal segmentation Error')
    return win, list(mean), list(std), float(shift[0]), float(scale[0])


# ---- lstmmodel.py -------------------------------------------------------------------------------
class _PredictModel:
    """Stands in for the Keras predict model returned by get_model1/get_model2 (lstmmodel.py:76-81)."""

    def __init__(self, which: int, reviser: engine.Reviser):
        self._which = which
        self._r = reviser

    def predict(self, inputs, batch_size=None, verbose=0):
        S, X = inputs
        p = self._r.predict_windows(np.asarray(S), np.asarray(X))
        return p[self._which]


def get_model1(SENT_LEN=None, reviser: Optional[engine.Reviser] = None):
    r = _rev(reviser)
    if SENT_LEN is not None and int(SENT_LEN) != r.window:
        raise ValueError("the loaded weights have window %d (feature.kernel.shape[0] // 6)" % r.window)
    return _PredictModel(0, r)


def get_model2(SENT_LEN=None, reviser: Optional[engine.Reviser] = None):
    r = _rev(reviser)
    if SENT_LEN is not None and int(SENT_LEN) != r.window:
        raise ValueError("the loaded weights have window %d (feature.kernel.shape[0] // 6)" % r.window)
    return _PredictModel(1, r)


# ---- output_handeler.py ---------------------------------------------------------------------------
label_to_base = {5: 'A', 4: 'G', 3: 'T', 2: 'C', 1: '-', 0: 'D'}


def get_base_1(event_bases, y_pre, y_pre2, reviser: Optional[engine.Reviser] = None):
    """``get_base_1(event_bases, y_pre, y_pre2)`` (output_handeler.py:104-122) on the GPU.

    ``y_pre`` are model-1 labels 0..5; ``y_pre2`` is what the reference function expects, i.e. the
    model-2 label + 1 (it subtracts 1 itself, :106) = ``argmax(P2) + 2``.  zip() truncation to the
    shortest input is reproduced.
    """
    r = _rev(reviser)
    y1 = np.asarray(y_pre).astype(np.int64)
    l2 = np.asarray(y_pre2).astype(np.int64) - 1
    M = min(len(event_bases), len(y1), len(l2))
    if len(y1) == 0:
        raise IndexError('index 0 is out of bounds')            # label_to_base[y_pre[0]] in the reference
    if np.any(l2[:M] < 1) or np.any(l2[:M] > 5) or np.any(y1[:M] < 0) or np.any(y1[:M] > 5):
        raise ValueError("labels outside the model's label space")
    bases = np.frombuffer(''.join(event_bases[:M]).encode('ascii'), dtype=np.uint8)
    W, bef = r.window, (r.window - 1) // 2
    aft = W - bef
    # embed as the core of a read with pass-through edges, then strip the edges again
    padded = np.concatenate([np.full(bef, ord('N'), np.uint8), bases, np.full(aft, ord('N'), np.uint8)])
    if M == 0:
        return label_to_base[int(y1[0])].replace('-', '')
    y1m = y1[:M].copy()
    rev, off = r.decode(np.array([0, M + W], np.int64), padded, y1m.astype(np.uint8), (l2[:M] - 1).astype(np.uint8))
    s = rev.tobytes().decode('ascii')
    return s[bef:len(s) - aft]


def prep_read_fasta(fast5_fn, read_fasta_fn, bases):
    try:
        fast5_fn = fast5_fn.split('/')[-1]
        text = ">" + fast5_fn.replace(' ', '|||') + '\n' + ''.join(bases)
        with open(read_fasta_fn, 'w') as fp:
            fp.write(text)
    except Exception:
        raise NotImplementedError('Error in writing .fasta file')
    return True


def prep_read_fastq(fast5_fn, read_fastq_fn, bases, qul):
    try:
        fast5_fn = fast5_fn.split('/')[-1]
        text = "@" + fast5_fn.replace(' ', '|||') + '\n' + ''.join(bases) + '+\n' + ''.join(qul)
        with open(read_fastq_fn, 'w') as fp:
            fp.write(text)
    except Exception:
        raise NotImplementedError('Error in writing .fastq file')
    return True


# ---- the NN path of provide_fasta, batched ------------------------------------------------------------
def revise_reads(reads: Sequence[ReadArrays], reviser: Optional[engine.Reviser] = None, want_labels=False,
                 want_probs=False, want_qual=False) -> engine.ReviseResult:
    """get_read_data outputs of many reads -> revised sequences (one ragged GPU batch); ``want_qual`` adds the
    fastq quality string of every revised read (``ReviseResult.quality(i)``, definition D6')."""
    r = _rev(reviser)
    return r.revise_batch(engine.pack_batch(reads), want_labels=want_labels, want_probs=want_probs, want_qual=want_qual)


def out_filename(output_dir: str, fast5_fn_sg: str, fmt: str) -> str:
    """NanoReviser.py:137,163 (string concatenation: ``-o`` must end in '/')."""
    return output_dir + fast5_fn_sg.split('.')[0] + '_out.' + fmt
