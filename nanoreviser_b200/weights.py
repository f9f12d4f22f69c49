"""Loader for the reference's per-species Keras-2.2.4 ``save_weights`` files.

The reference derives the two weight paths from ``-S`` at ``NanoReviser.py:192-193``
(``./model/<S>/<S>_win13_50ep_model{1,2}.h5``) and the graph those files were saved
from is ``nanorevutils/lstmmodel.py:32-81`` (model1) / ``:84-133`` (model2).

Layers are mapped by *position and role* in the root ``layer_names`` attribute, not
by name: model2 files carry other numeric suffixes (``time_distributed_19``,
``bidirectional_13``, ``final_out_3`` ...).  The window length is read from the
weights (``feature.kernel.shape[0] // 6`` = 11, although the files are named
``win13``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

from . import h5mini

SIGNAL_LEN = 50   # lstmmodel.py:28
VEC_LEN = 6       # lstmmodel.py:29
CNN_CH = 8        # lstmmodel.py:35  identity_Block(signal_input, 8, 3)
LSTM_UNITS = (16, 64, 128, 64)      # lstmmodel.py:44,46,49,51
LSTM_INPUTS = (6, 32, 192, 256)


@dataclass
class LstmDir:
    kernel: np.ndarray      # (in, 4u)   gate order i,f,c,o (Keras 2.2.4)
    recurrent: np.ndarray   # (u, 4u)
    bias: np.ndarray        # (4u,)


@dataclass
class ModelWeights:
    window: int
    n_class: int
    conv1_k: np.ndarray     # (3,1,8)
    conv1_b: np.ndarray
    bn1: np.ndarray         # (4,8): gamma, beta, moving_mean, moving_variance
    conv2_k: np.ndarray     # (3,8,8)
    conv2_b: np.ndarray
    bn2: np.ndarray
    sig_dense_k: np.ndarray  # (400,64)
    sig_dense_b: np.ndarray
    lstm: List[Tuple[LstmDir, LstmDir]] = field(default_factory=list)  # 4 x (fwd, bwd)
    bn_rnn: List[np.ndarray] = field(default_factory=list)             # after lstm 0,1,2: (4,32),(4,128),(4,256)
    dense1_k: np.ndarray = None   # (128,128)
    dense1_b: np.ndarray = None
    dense2_k: np.ndarray = None   # (128,32)
    dense2_b: np.ndarray = None
    main_k: np.ndarray = None     # (32,6)
    main_b: np.ndarray = None
    feat_k: np.ndarray = None     # (W*6,16)
    feat_b: np.ndarray = None
    final_k: np.ndarray = None    # (16,n_class)
    final_b: np.ndarray = None

    def n_params(self) -> int:
        n = 0
        for v in self.__dict__.values():
            if isinstance(v, np.ndarray):
                n += v.size
        for f, b in self.lstm:
            for d in (f, b):
                n += d.kernel.size + d.recurrent.size + d.bias.size
        for bn in self.bn_rnn:
            n += bn.size
        return n


def _read_layers(fn: str):
    """-> list of (layer_name, [arrays in weight_names order]) for layers that own weights."""
    out = []
    with h5mini.File(fn) as f:
        for ln in f.attrs["layer_names"]:
            ln = ln.decode() if isinstance(ln, bytes) else str(ln)
            g = f[ln]
            wn = g.attrs.get("weight_names")
            if wn is None or len(wn) == 0:
                continue
            arrs = []
            for w in wn:
                w = w.decode() if isinstance(w, bytes) else str(w)
                arrs.append(np.ascontiguousarray(g[w][()], dtype=np.float32))
            out.append((ln, arrs))
    return out


def _expect(cond, msg):
    if not cond:
        raise ValueError("unexpected weight file layout: " + msg)


def load_model_weights(fn: str) -> ModelWeights:
    layers = _read_layers(fn)
    _expect(len(layers) == 17, "%d weighted layers (want 17)" % len(layers))
    it = iter(layers)

    def nxt(n_arrays, what):
        name, arrs = next(it)
        _expect(len(arrs) == n_arrays, "%s: layer %s has %d arrays" % (what, name, len(arrs)))
        return arrs

    conv1_k, conv1_b = nxt(2, "conv1")
    bn1 = np.stack(nxt(4, "bn1"))
    conv2_k, conv2_b = nxt(2, "conv2")
    bn2 = np.stack(nxt(4, "bn2"))
    _expect(conv1_k.shape == (3, 1, CNN_CH) and conv2_k.shape == (3, CNN_CH, CNN_CH), "conv kernels")
    lstm, bn_rnn = [], []

    def bidir(idx):
        a = nxt(6, "bidirectional %d" % idx)
        u, i = LSTM_UNITS[idx], LSTM_INPUTS[idx]
        for k in (0, 3):
            _expect(a[k].shape == (i, 4 * u) and a[k + 1].shape == (u, 4 * u) and a[k + 2].shape == (4 * u,),
                    "lstm %d shapes" % idx)
        lstm.append((LstmDir(a[0], a[1], a[2]), LstmDir(a[3], a[4], a[5])))

    bidir(0)
    bn_rnn.append(np.stack(nxt(4, "bn after read_rnn1")))
    bidir(1)
    bn_rnn.append(np.stack(nxt(4, "bn after read_rnn11")))
    sig_dense_k, sig_dense_b = nxt(2, "signal dense")
    _expect(sig_dense_k.shape == (SIGNAL_LEN * CNN_CH, 64), "signal dense kernel")
    bidir(2)
    bn_rnn.append(np.stack(nxt(4, "bn after total_rnn1")))
    bidir(3)
    _expect([b.shape for b in bn_rnn] == [(4, 32), (4, 128), (4, 256)], "batch-norm shapes")
    dense1_k, dense1_b = nxt(2, "dense1")
    dense2_k, dense2_b = nxt(2, "dense2")
    main_k, main_b = nxt(2, "main_out")
    feat_k, feat_b = nxt(2, "feature")
    final_k, final_b = nxt(2, "final_out")
    _expect(dense1_k.shape == (128, 128) and dense2_k.shape == (128, 32) and main_k.shape == (32, 6), "dense heads")
    _expect(feat_k.shape[0] % 6 == 0 and feat_k.shape[1] == 16 and final_k.shape[0] == 16, "feature/final")
    window = feat_k.shape[0] // 6
    n_class = final_b.shape[0]
    _expect(n_class in (5, 6), "n_class %d" % n_class)
    return ModelWeights(window=window, n_class=n_class, conv1_k=conv1_k, conv1_b=conv1_b, bn1=bn1,
                        conv2_k=conv2_k, conv2_b=conv2_b, bn2=bn2, sig_dense_k=sig_dense_k,
                        sig_dense_b=sig_dense_b, lstm=lstm, bn_rnn=bn_rnn, dense1_k=dense1_k,
                        dense1_b=dense1_b, dense2_k=dense2_k, dense2_b=dense2_b, main_k=main_k,
                        main_b=main_b, feat_k=feat_k, feat_b=feat_b, final_k=final_k, final_b=final_b)


def model_paths(species: str, root: str = "./model"):
    """``NanoReviser.py:192-193``."""
    s = str(species)
    return (root.rstrip("/") + "/" + s + "/" + s + "_win13_50ep_model1.h5",
            root.rstrip("/") + "/" + s + "/" + s + "_win13_50ep_model2.h5")


def load_species(species: str, root: str = "./model"):
    p1, p2 = model_paths(species, root)
    m1, m2 = load_model_weights(p1), load_model_weights(p2)
    if m1.n_class != 6 or m2.n_class != 5:
        raise ValueError("model1 must have 6 classes and model2 5 (got %d/%d)" % (m1.n_class, m2.n_class))
    if m1.window != m2.window:
        raise ValueError("model1/model2 window mismatch")
    return m1, m2
