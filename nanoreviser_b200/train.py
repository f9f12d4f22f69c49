"""Training of the two NanoReviser networks on the GPU.

The reference trains with Keras (NanoReviser_train.py:164-199):

    att_model_train, att_model_predict = get_model1(window)                       # lstmmodel.py:32-81
    att_model_train.fit([signal_x[..., None], x, y], [y, zeros], class_weight=..., validation_split=..., shuffle=True,
                        epochs=..., batch_size=..., verbose=...)
    att_model_predict.save_weights(fn)

:class:`TrainModel` is that object for this path: ``fit`` has Model.fit's arguments and returns the ``history`` dictionary
(loss, final_out_loss, l2_loss1_loss, final_out_acc, l2_loss1_acc and their val_ twins), ``save_weights`` writes the predict model's weights
in the Keras 2.2.4 HDF5 layout the inference path (weights.load_model_weights -> engine.Reviser) and Keras itself read.

Every arithmetic step of a training step runs in the hand-written CUDA operators of csrc/nrv_train.cu (include/nrv_train.h):
convolutions, batch normalisation with batch statistics, the four Bi-LSTMs step by step with full back-propagation through
time, the dense heads, softmax cross-entropy + centre loss (weights 1.0 / 0.4, lstmmodel.py:69-74) and Adam.  This module only
decides which buffer goes into which operator; PyTorch is used for device memory and the stream, nothing else.  Without a CUDA
device, or without the library, construction fails -- there is no CPU path.

Semantics restated from the Keras 2.2.4 / TensorFlow 1.12 sources, NOT pinned against a Keras run (neither is installable
here); the gradients themselves are pinned against an fp64 autograd graph of the same network
(tests/test_train_gpu.py).  Choices that a Keras run would have to confirm:
  * BatchNormalization: batch statistics with the biased variance, epsilon 1e-3, moving averages with momentum 0.99 where the
    moving variance takes the batch variance times n / (n - 1 - epsilon) (keras/layers/normalization.py:  sample_size correction);
  * Dropout(0.2): inverted dropout, an independent mask per element and step (a counter-based generator of our own, nrvt_dropout_mask, not TensorFlow's);
  * class_weight: Keras 2.2.4 interprets a flat ``{class: weight}`` dictionary on a model with two outputs as
    ``{output_name: ...}`` and therefore applies NO class weights to either output; ``class_weight_mode='keras'`` (default)
    reproduces that, ``'applied'`` weights the cross-entropy samples the way the reference's author evidently intended;
  * Adam: lr 1e-3, beta 0.9 / 0.999, epsilon 1e-7, the bias correction folded into the step size (keras/optimizers.py Adam).
Runs with the same seeds agree to fp32 rounding (split-K products accumulate with atomics); NRV_TRAIN_DETERMINISTIC=1 makes them
bit-identical at about a third of the speed.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import time
from typing import Dict, List, Optional

import numpy as np

from . import engine, h5write
from .weights import CNN_CH, LSTM_INPUTS, LSTM_UNITS, SIGNAL_LEN, LstmDir, ModelWeights

BN_EPS = 1e-3
BN_MOMENTUM = 0.99
DROPOUT = 0.2
CENTER_DIM = 16
LOSS_WEIGHTS = (1.0, 0.4)          # lstmmodel.py:71


class TrainError(RuntimeError):
    pass


def _lib():
    lib = engine.load_library()
    if getattr(lib, "_nrvt_ready", False):
        return lib
    f32p, vp, i64 = C.c_void_p, C.c_void_p, C.c_int64
    sig = {
        "nrvt_gemm": [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, f32p, C.c_int, f32p, C.c_int, C.c_float, f32p, C.c_int],
        "nrvt_bias_act": [vp, f32p, C.c_int, C.c_int, C.c_int, f32p, C.c_int],
        "nrvt_relu_bwd": [vp, f32p, f32p, i64],
        "nrvt_colsum": [vp, f32p, C.c_int, C.c_int, C.c_int, f32p, C.c_float],
        "nrvt_copy2d": [vp, f32p, C.c_int, f32p, C.c_int, C.c_int, C.c_int, C.c_int],
        "nrvt_conv1d_fwd": [vp, f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int, C.c_int],
        "nrvt_conv1d_bwd": [vp, f32p, f32p, f32p, f32p, f32p, f32p, f32p, vp, C.c_int, C.c_int, C.c_int, C.c_int],
        "nrvt_add_bcast": [vp, f32p, f32p, i64, C.c_int],
        "nrvt_bn_fwd": [vp, f32p, f32p, f32p, C.c_float, f32p, f32p, f32p, vp, i64, C.c_int],
        "nrvt_bn_apply": [vp, f32p, f32p, f32p, f32p, f32p, C.c_float, f32p, i64, C.c_int],
        "nrvt_bn_bwd": [vp, f32p, f32p, f32p, f32p, f32p, C.c_float, f32p, f32p, f32p, vp, i64, C.c_int],
        "nrvt_lstm_cell_fwd": [vp, f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_int],
        "nrvt_lstm_cell_bwd": [vp, f32p, f32p, f32p, f32p, C.c_int, f32p, f32p, f32p, C.c_int, C.c_int],
        "nrvt_softmax_ce": [vp, f32p, vp, f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_float],
        "nrvt_center_loss": [vp, f32p, vp, f32p, f32p, f32p, f32p, C.c_int, C.c_int, C.c_float],
        "nrvt_dropout_mask": [vp, vp, i64, C.c_uint64, vp, C.c_float],
        "nrvt_dropout": [vp, f32p, vp, i64, C.c_float],
        "nrvt_adam": [vp, f32p, f32p, f32p, f32p, i64, f32p, C.c_float, C.c_float, C.c_float],
        "nrvt_ema": [vp, f32p, f32p, i64, C.c_float, C.c_float],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes, fn.restype = args, C.c_int
    lib.nrvt_last_error.restype = C.c_char_p
    lib._nrvt_ready = True
    return lib


TRAIN_EXPORTS = ["nrvt_last_error", "nrvt_gemm", "nrvt_bias_act", "nrvt_relu_bwd", "nrvt_colsum", "nrvt_copy2d", "nrvt_conv1d_fwd",
                 "nrvt_conv1d_bwd", "nrvt_add_bcast", "nrvt_bn_fwd", "nrvt_bn_apply", "nrvt_bn_bwd", "nrvt_lstm_cell_fwd", "nrvt_lstm_cell_bwd",
                 "nrvt_softmax_ce", "nrvt_center_loss", "nrvt_dropout_mask", "nrvt_dropout", "nrvt_adam", "nrvt_ema"]


# -------------------------------------------------------------------------------------------------------------------------
# Keras default initialisers (keras/initializers.py) for a model trained from scratch
# -------------------------------------------------------------------------------------------------------------------------
def _glorot_uniform(rng, shape, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def _orthogonal(rng, rows, cols):
    a = rng.normal(0.0, 1.0, (rows, cols))
    u, _, vt = np.linalg.svd(a, full_matrices=False)
    q = u if u.shape == (rows, cols) else vt
    return q.astype(np.float32)


def init_weights(window: int, n_class: int, seed: int = 0) -> ModelWeights:
    """glorot_uniform kernels, orthogonal recurrent kernels, zero biases with the forget-gate bias at 1 (unit_forget_bias),
    BatchNormalization gamma 1 / beta 0 / moving mean 0 / moving variance 1."""
    rng = np.random.default_rng(seed)
    bn = lambda c: np.stack([np.ones(c), np.zeros(c), np.zeros(c), np.ones(c)]).astype(np.float32)
    dense = lambda i, o: (_glorot_uniform(rng, (i, o), i, o), np.zeros(o, np.float32))
    lstm = []
    for u, i in zip(LSTM_UNITS, LSTM_INPUTS):
        dirs = []
        for _ in range(2):
            b = np.zeros(4 * u, np.float32)
            b[u:2 * u] = 1.0
            dirs.append(LstmDir(_glorot_uniform(rng, (i, 4 * u), i, 4 * u), _orthogonal(rng, u, 4 * u), b))
        lstm.append(tuple(dirs))
    d1k, d1b = dense(128, 128); d2k, d2b = dense(128, 32); mk, mb = dense(32, 6)
    fk, fb = dense(window * 6, CENTER_DIM); ok, ob = dense(CENTER_DIM, n_class)
    sk, sb = dense(SIGNAL_LEN * CNN_CH, 64)
    return ModelWeights(window=window, n_class=n_class,
                        conv1_k=_glorot_uniform(rng, (3, 1, CNN_CH), 3, 3 * CNN_CH), conv1_b=np.zeros(CNN_CH, np.float32), bn1=bn(CNN_CH),
                        conv2_k=_glorot_uniform(rng, (3, CNN_CH, CNN_CH), 3 * CNN_CH, 3 * CNN_CH), conv2_b=np.zeros(CNN_CH, np.float32),
                        bn2=bn(CNN_CH), sig_dense_k=sk, sig_dense_b=sb, lstm=lstm, bn_rnn=[bn(32), bn(128), bn(256)],
                        dense1_k=d1k, dense1_b=d1b, dense2_k=d2k, dense2_b=d2b, main_k=mk, main_b=mb, feat_k=fk, feat_b=fb,
                        final_k=ok, final_b=ob)


# -------------------------------------------------------------------------------------------------------------------------
class TrainModel:
    """att_model_train / att_model_predict of get_model1 (n_class 6) / get_model2 (n_class 5) as one object."""

    def __init__(self, window: int = 13, n_class: int = 6, weights: Optional[ModelWeights] = None, device: int = 0, seed: int = 0,
                 class_weight_mode: str = "keras"):
        import torch
        self._torch = torch
        if not torch.cuda.is_available():
            raise TrainError("no CUDA device: the training path has no CPU implementation")
        if class_weight_mode not in ("keras", "applied"):
            raise ValueError("class_weight_mode must be 'keras' or 'applied'")
        self.lib = _lib()
        self.dev = torch.device("cuda", device)
        self.window, self.n_class = int(window), int(n_class)
        self.class_weight_mode = class_weight_mode
        w = weights if weights is not None else init_weights(self.window, self.n_class, seed)
        if w.window != self.window or w.n_class != self.n_class:
            raise ValueError("weights are for window %d / %d classes" % (w.window, w.n_class))
        self.p: Dict[str, "torch.Tensor"] = {}        # trainable parameters
        self.s: Dict[str, "torch.Tensor"] = {}        # BatchNormalization moving statistics
        put = lambda k, a: self.p.__setitem__(k, torch.tensor(np.ascontiguousarray(a, np.float32), device=self.dev))
        put("conv1_k", w.conv1_k); put("conv1_b", w.conv1_b); put("conv2_k", w.conv2_k); put("conv2_b", w.conv2_b)
        put("sig_k", w.sig_dense_k); put("sig_b", w.sig_dense_b)
        for name, bn in (("bn1", w.bn1), ("bn2", w.bn2), ("bnr0", w.bn_rnn[0]), ("bnr1", w.bn_rnn[1]), ("bnr2", w.bn_rnn[2])):
            put(name + "_g", bn[0]); put(name + "_b", bn[1])
            self.s[name + "_mean"] = torch.tensor(bn[2].astype(np.float32), device=self.dev)
            self.s[name + "_var"] = torch.tensor(bn[3].astype(np.float32), device=self.dev)
        for l, pair in enumerate(w.lstm):
            for d, ld in enumerate(pair):
                put("l%d%d_k" % (l, d), ld.kernel); put("l%d%d_r" % (l, d), ld.recurrent); put("l%d%d_b" % (l, d), ld.bias)
        put("d1_k", w.dense1_k); put("d1_b", w.dense1_b); put("d2_k", w.dense2_k); put("d2_b", w.dense2_b)
        put("m_k", w.main_k); put("m_b", w.main_b); put("f_k", w.feat_k); put("f_b", w.feat_b)
        put("o_k", w.final_k); put("o_b", w.final_b)
        rng = np.random.default_rng(seed + 1)
        put("centers", rng.uniform(-0.05, 0.05, (self.n_class, CENTER_DIM)))      # Embedding default initialiser
        self.g = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.p.items()}
        self.iterations = 0
        self.lr, self.beta1, self.beta2, self.eps = 1e-3, 0.9, 0.999, 1e-7
        self._buf: Dict[str, "torch.Tensor"] = {}
        self._pool: Dict[tuple, "torch.Tensor"] = {}
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        # per-step scalars live on the device (the step is captured into a CUDA graph, where by-value arguments would freeze)
        self._step_d = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._lr_t_d = torch.zeros(1, dtype=torch.float32, device=self.dev)
        self._graphs: Dict[tuple, dict] = {}
        self.launches = 0
        self.use_graph = os.environ.get("NRV_TRAIN_GRAPH", "1") != "0"      # fit(): steps replayed from a CUDA graph

    # -- plumbing ------------------------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(self._torch.cuda.current_stream(self.dev).cuda_stream)

    def _call(self, name, *args):
        rc = getattr(self.lib, name)(self._stream(), *args)
        self.launches += 1
        if rc != 0:
            raise TrainError("%s failed (%d): %s" % (name, rc, self.lib.nrvt_last_error().decode()))

    def buf(self, name, shape, dtype=None, zero=False):
        """Named scratch tensor.  One tensor per (name, shape, dtype) is kept for the life of the model -- a captured CUDA graph holds
        raw pointers into the buffers of its batch size, which therefore must survive passes with another batch size -- and
        self._buf[name] is the one used last."""
        torch = self._torch
        dtype = dtype or torch.float32
        key = (name, tuple(int(v) for v in shape), dtype)
        t = self._pool.get(key)
        if t is None:
            t = torch.empty(key[1], dtype=dtype, device=self.dev)
            self._pool[key] = t
        self._buf[name] = t
        if zero:
            t.zero_()
        return t

    @staticmethod
    def _ptr(t, offset_elems: int = 0):
        return C.c_void_p(t.data_ptr() + 4 * int(offset_elems)) if t is not None else None

    def gemm(self, ta, tb, M, N, K, A, lda, B, ldb, Cm, ldc, beta=0.0, alpha=1.0, ao=0, bo=0, co=0):
        self._call("nrvt_gemm", ta, tb, M, N, K, alpha, self._ptr(A, ao), lda, self._ptr(B, bo), ldb, beta, self._ptr(Cm, co), ldc)

    # -- layers ----------------------------------------------------------------------------------------------------------------
    def _bn_fwd(self, name, X, rows, Cn, training):
        """training: batch statistics (kept for the backward pass and the moving-average update); else the moving statistics"""
        Y = self.buf(name + "_y", (rows, Cn))
        if training:
            mean, var = self.buf(name + "_bm", (Cn,)), self.buf(name + "_bv", (Cn,))
            work = self.buf("bn_work", (4 * 256,), dtype=self._torch.float64)
            self._call("nrvt_bn_fwd", self._ptr(X), self._ptr(self.p[name + "_g"]), self._ptr(self.p[name + "_b"]), BN_EPS,
                       self._ptr(Y), self._ptr(mean), self._ptr(var), C.c_void_p(work.data_ptr()), rows, Cn)
        else:
            self._call("nrvt_bn_apply", self._ptr(X), self._ptr(self.p[name + "_g"]), self._ptr(self.p[name + "_b"]),
                       self._ptr(self.s[name + "_mean"]), self._ptr(self.s[name + "_var"]), BN_EPS, self._ptr(Y), rows, Cn)
        return Y

    def _bn_bwd(self, name, X, dY, rows, Cn):
        dX = self.buf(name + "_dx", (rows, Cn))
        work = self.buf("bn_work", (4 * 256,), dtype=self._torch.float64)
        self._call("nrvt_bn_bwd", self._ptr(X), self._ptr(dY), self._ptr(self.p[name + "_g"]), self._ptr(self.buf(name + "_bm", (Cn,))),
                   self._ptr(self.buf(name + "_bv", (Cn,))), BN_EPS, self._ptr(dX), self._ptr(self.g[name + "_g"]),
                   self._ptr(self.g[name + "_b"]), C.c_void_p(work.data_ptr()), rows, Cn)
        return dX

    def _lstm_fwd(self, l, xin, ld_in, n_in, T, B):
        u = LSTM_UNITS[l]
        n = T * B
        out = self.buf("out%d" % l, (n, 2 * u))
        for d in range(2):
            z = self.buf("z%d%d" % (l, d), (n, 4 * u))
            c = self.buf("c%d%d" % (l, d), (n, u))
            self.gemm(0, 0, n, 4 * u, n_in, xin, ld_in, self.p["l%d%d_k" % (l, d)], 4 * u, z, 4 * u)
            self._call("nrvt_bias_act", self._ptr(z), n, 4 * u, 4 * u, self._ptr(self.p["l%d%d_b" % (l, d)]), 0)
            prev = None
            for s in range(T):
                t = s if d == 0 else T - 1 - s
                if prev is not None:        # z_t += h_{prev} . Wr
                    self.gemm(0, 0, B, 4 * u, u, out, 2 * u, self.p["l%d%d_r" % (l, d)], 4 * u, z, 4 * u, beta=1.0,
                              ao=prev * B * 2 * u + d * u, co=t * B * 4 * u)
                self._call("nrvt_lstm_cell_fwd", self._ptr(z, t * B * 4 * u), self._ptr(c, prev * B * u) if prev is not None else None,
                           self._ptr(c, t * B * u), self._ptr(out, t * B * 2 * u + d * u), 2 * u, B, u)
                prev = t
        return out

    def _lstm_bwd(self, l, xin, ld_in, n_in, dout, T, B, want_dx=True):
        """dout [n, 2u] -> gradients of the layer's parameters (into self.g) and, if wanted, d(xin) [n, n_in]"""
        u = LSTM_UNITS[l]
        n = T * B
        out = self._buf["out%d" % l]
        dx = self.buf("dx%d" % l, (n, n_in)) if want_dx else None
        for d in range(2):
            z, c = self._buf["z%d%d" % (l, d)], self._buf["c%d%d" % (l, d)]
            dz = self.buf("dz%d" % l, (n, 4 * u))
            dc = self.buf("dc%d" % l, (B, u), zero=True)
            dh = self.buf("dh%d" % l, (B, u))
            Wr, Wk = self.p["l%d%d_r" % (l, d)], self.p["l%d%d_k" % (l, d)]
            for s in reversed(range(T)):
                t = s if d == 0 else T - 1 - s
                prev = None if s == 0 else (s - 1 if d == 0 else T - s)
                self._call("nrvt_lstm_cell_bwd", self._ptr(z, t * B * 4 * u), self._ptr(c, prev * B * u) if prev is not None else None,
                           self._ptr(c, t * B * u), self._ptr(dout, t * B * 2 * u + d * u), 2 * u,
                           self._ptr(dh) if s != T - 1 else None, self._ptr(dc), self._ptr(dz, t * B * 4 * u), B, u)
                if s > 0:                   # dh_{prev} = dz_t . Wr^T
                    self.gemm(0, 1, B, u, 4 * u, dz, 4 * u, Wr, 4 * u, dh, u, ao=t * B * 4 * u)
            # parameter gradients: dWk = X^T dZ, db = colsum dZ, dWr = sum_t h_prev(t)^T dz_t (one GEMM over the shifted time range)
            self.gemm(1, 0, n_in, 4 * u, n, xin, ld_in, dz, 4 * u, self.g["l%d%d_k" % (l, d)], 4 * u)
            self._call("nrvt_colsum", self._ptr(dz), n, 4 * u, 4 * u, self._ptr(self.g["l%d%d_b" % (l, d)]), 0.0)
            if T > 1:
                h_off = (0 if d == 0 else B * 2 * u) + d * u            # fwd: h_prev(t) = out[t-1]; bwd: out[t+1]
                z_off = B * 4 * u if d == 0 else 0
                self.gemm(1, 0, u, 4 * u, (T - 1) * B, out, 2 * u, dz, 4 * u, self.g["l%d%d_r" % (l, d)], 4 * u, ao=h_off, bo=z_off)
            else:
                self.g["l%d%d_r" % (l, d)].zero_()
            if want_dx:
                self.gemm(0, 1, n, n_in, 4 * u, dz, 4 * u, Wk, 4 * u, dx, n_in, beta=0.0 if d == 0 else 1.0)
        return dx

    def _dense_fwd(self, name, X, ldx, rows, n_in, n_out, relu, out_name):
        Y = self.buf(out_name, (rows, n_out))
        self.gemm(0, 0, rows, n_out, n_in, X, ldx, self.p[name + "_k"], n_out, Y, n_out)
        self._call("nrvt_bias_act", self._ptr(Y), rows, n_out, n_out, self._ptr(self.p[name + "_b"]), int(relu))
        return Y

    def _dense_bwd(self, name, X, ldx, Y, dY, rows, n_in, n_out, relu, dx_name):
        """dY is consumed (masked in place when relu); returns dX [rows, n_in] or None"""
        if relu:
            self._call("nrvt_relu_bwd", self._ptr(dY), self._ptr(Y), rows * n_out)
        self.gemm(1, 0, n_in, n_out, rows, X, ldx, dY, n_out, self.g[name + "_k"], n_out)
        self._call("nrvt_colsum", self._ptr(dY), rows, n_out, n_out, self._ptr(self.g[name + "_b"]), 0.0)
        if dx_name is None:
            return None
        dX = self.buf(dx_name, (rows, n_in))
        self.gemm(0, 1, rows, n_in, n_out, dY, n_out, self.p[name + "_k"], n_out, dX, n_in)
        return dX

    # -- one batch ---------------------------------------------------------------------------------------------------------------
    def _upload(self, S, X, y):
        """numpy [B,T,50] / [B,T,6] / [B] -> time-major device tensors"""
        torch = self._torch
        S = np.asarray(S, np.float32)
        if S.ndim == 4:
            S = S[..., 0]
        X = np.asarray(X, np.float32)
        B, T = X.shape[0], X.shape[1]
        if T != self.window or S.shape != (B, T, SIGNAL_LEN) or X.shape != (B, T, 6):
            raise ValueError("expected signal [B,%d,50(,1)] and read [B,%d,6]" % (self.window, self.window))
        y = np.asarray(y).reshape(-1)
        if y.shape[0] != B:
            raise ValueError("one label per window")
        yi = y.astype(np.int64)
        if (yi < 0).any() or (yi >= self.n_class).any():
            raise ValueError("labels must be in [0, %d)" % self.n_class)
        S_tm = torch.from_numpy(np.ascontiguousarray(S.transpose(1, 0, 2))).to(self.dev, non_blocking=True)
        X_tm = torch.from_numpy(np.ascontiguousarray(X.transpose(1, 0, 2))).to(self.dev, non_blocking=True)
        y_d = torch.from_numpy(yi.astype(np.int32)).to(self.dev, non_blocking=True)
        return S_tm, X_tm, y_d, B, T

    def forward_backward(self, S, X, y, class_weight=None, training=True, dropout_mask=None):
        """One batch through the train model.  training=True: batch statistics, dropout, gradients left in self.g.
        Returns {'loss', 'final_out_loss', 'l2_loss1_loss', 'final_out_acc', 'l2_loss1_acc'} (device-side sums turned into batch means)."""
        S_tm, X_tm, y_d, B, T = self._upload(S, X, y)
        if not training:
            return self._evaluate(S_tm, X_tm, y_d, B, T)
        mask = None
        if dropout_mask is not None:        # [B,T,50,8] booleans (tests): to the time-major order
            mask = self._torch.from_numpy(np.ascontiguousarray(np.asarray(dropout_mask).transpose(1, 0, 2, 3)).astype(np.uint8)).to(self.dev)
        self._step_d.fill_(self.iterations)
        return self._metrics(self._train_device(S_tm, X_tm, y_d, B, T, self._class_weight_tensor(class_weight), mask), B)

    def _class_weight_tensor(self, class_weight):
        if class_weight is None or self.class_weight_mode != "applied":
            return None
        return self._torch.tensor([float(class_weight.get(k, 1.0)) for k in range(self.n_class)], dtype=self._torch.float32, device=self.dev)

    def _forward(self, S_tm, X_tm, B, T, training, mask=None):
        """The graph of lstmmodel.py:32-63 on time-major inputs.  training: batch statistics + dropout (train model), else moving
        statistics, no dropout (predict model).  Returns the tensors the losses and the backward pass need."""
        torch = self._torch
        n = T * B
        L, Cc = SIGNAL_LEN, CNN_CH
        # ---- CNN branch (nanorevcnn.py:29-38, lstmmodel.py:35-41) on the n = T*B signals
        c1 = self.buf("c1", (n * L, Cc))
        self._call("nrvt_conv1d_fwd", self._ptr(S_tm), self._ptr(self.p["conv1_k"]), self._ptr(self.p["conv1_b"]), self._ptr(c1), n, L, 1, Cc)
        b1 = self._bn_fwd("bn1", c1, n * L, Cc, training)
        c2 = self.buf("c2", (n * L, Cc))
        self._call("nrvt_conv1d_fwd", self._ptr(b1), self._ptr(self.p["conv2_k"]), self._ptr(self.p["conv2_b"]), self._ptr(c2), n, L, Cc, Cc)
        res = self._bn_fwd("bn2", c2, n * L, Cc, training)
        self._call("nrvt_add_bcast", self._ptr(res), self._ptr(S_tm), n * L, Cc)
        if training:
            if mask is None:
                mask = self.buf("mask", (n * L, Cc), dtype=torch.uint8)
                self._call("nrvt_dropout_mask", C.c_void_p(mask.data_ptr()), n * L * Cc, self.seed, C.c_void_p(self._step_d.data_ptr()), DROPOUT)
            self._call("nrvt_dropout", self._ptr(res), C.c_void_p(mask.data_ptr()), n * L * Cc, 1.0 / (1.0 - DROPOUT))
        # ---- read branch: Bi-LSTM(16) -> BN -> Bi-LSTM(64) -> BN ; concat [read_rnn2 (128) | signal (64)]
        out0 = self._lstm_fwd(0, X_tm, 6, 6, T, B)
        r1 = self._bn_fwd("bnr0", out0, n, 32, training)
        out1 = self._lstm_fwd(1, r1, 32, 32, T, B)
        r2 = self._bn_fwd("bnr1", out1, n, 128, training)
        tot = self.buf("tot", (n, 192))
        self._call("nrvt_copy2d", self._ptr(tot), 192, self._ptr(r2), 128, n, 128, 0)
        self.gemm(0, 0, n, 64, L * Cc, res, L * Cc, self.p["sig_k"], 64, tot, 192, co=128)       # TD(Dense 64) straight into the concat
        self._call("nrvt_bias_act", self._ptr(tot, 128), n, 64, 192, self._ptr(self.p["sig_b"]), 0)
        out2 = self._lstm_fwd(2, tot, 192, 192, T, B)
        t1 = self._bn_fwd("bnr2", out2, n, 256, training)
        out3 = self._lstm_fwd(3, t1, 256, 256, T, B)
        # ---- heads (lstmmodel.py:55-63)
        d1 = self._dense_fwd("d1", out3, 128, n, 128, 128, True, "d1")
        d2 = self._dense_fwd("d2", d1, 128, n, 128, 32, True, "d2")
        d3 = self._dense_fwd("m", d2, 32, n, 32, 6, True, "d3")
        flat = self.buf("flat", (B, T * 6))                  # Flatten: index t*6 + k
        for t in range(T):
            self._call("nrvt_copy2d", self._ptr(flat, t * 6), T * 6, self._ptr(d3, t * B * 6), 6, B, 6, 0)
        feat = self._dense_fwd("f", flat, T * 6, B, T * 6, CENTER_DIM, True, "feat")
        logits = self._dense_fwd("o", feat, CENTER_DIM, B, CENTER_DIM, self.n_class, False, "logits")
        return dict(c1=c1, b1=b1, c2=c2, res=res, mask=mask, out0=out0, r1=r1, out1=out1, tot=tot, out2=out2, t1=t1, out3=out3,
                    d1=d1, d2=d2, d3=d3, flat=flat, feat=feat, logits=logits)

    def _train_device(self, S_tm, X_tm, y_d, B, T, cw, mask):
        """Forward + backward of one batch, device work only (no allocation once the buffers exist, no synchronisation): this is
        what a CUDA graph captures.  Returns the device tensor of the loss sums."""
        torch = self._torch
        n = T * B
        L, Cc = SIGNAL_LEN, CNN_CH
        keep_scale = 1.0 / (1.0 - DROPOUT)
        f = self._forward(S_tm, X_tm, B, T, True, mask)
        c1, b1, c2, res, mask, out0, r1, out1, tot, out2, t1, out3 = (f[k] for k in ("c1", "b1", "c2", "res", "mask", "out0", "r1", "out1", "tot",
                                                                                     "out2", "t1", "out3"))
        d1, d2, d3, flat, feat, logits = (f[k] for k in ("d1", "d2", "d3", "flat", "feat", "logits"))
        # ---- losses: 1.0 * mean(w_y ce) + 0.4 * mean(l2)
        stats = self.buf("stats", (4,), zero=True)
        probs = self.buf("probs", (B, self.n_class))
        dlog = self.buf("dlogits", (B, self.n_class))
        self._call("nrvt_softmax_ce", self._ptr(logits), C.c_void_p(y_d.data_ptr()), self._ptr(cw), self._ptr(probs), self._ptr(dlog),
                   self._ptr(stats), B, self.n_class, LOSS_WEIGHTS[0] / B)
        # ---- backward
        dfeat = self._dense_bwd("o", feat, CENTER_DIM, logits, dlog, B, CENTER_DIM, self.n_class, False, "dfeat")
        self.g["centers"].zero_()
        self._call("nrvt_center_loss", self._ptr(feat), C.c_void_p(y_d.data_ptr()), self._ptr(self.p["centers"]), self._ptr(dfeat),
                   self._ptr(self.g["centers"]), self._ptr(stats), B, CENTER_DIM, LOSS_WEIGHTS[1] / B)
        dflat = self._dense_bwd("f", flat, T * 6, feat, dfeat, B, T * 6, CENTER_DIM, True, "dflat")
        dd3 = self.buf("dd3", (n, 6))
        for t in range(T):
            self._call("nrvt_copy2d", self._ptr(dd3, t * B * 6), 6, self._ptr(dflat, t * 6), T * 6, B, 6, 0)
        dd2 = self._dense_bwd("m", d2, 32, d3, dd3, n, 32, 6, True, "dd2")
        dd1 = self._dense_bwd("d2", d1, 128, d2, dd2, n, 128, 32, True, "dd1")
        dout3 = self._dense_bwd("d1", out3, 128, d1, dd1, n, 128, 128, True, "dout3")
        dt1 = self._lstm_bwd(3, t1, 256, 256, dout3, T, B)
        dout2 = self._bn_bwd("bnr2", out2, dt1, n, 256)
        dtot = self._lstm_bwd(2, tot, 192, 192, dout2, T, B)
        dr2 = self.buf("dr2", (n, 128))
        self._call("nrvt_copy2d", self._ptr(dr2), 128, self._ptr(dtot), 192, n, 128, 0)
        dout1 = self._bn_bwd("bnr1", out1, dr2, n, 128)
        dr1 = self._lstm_bwd(1, r1, 32, 32, dout1, T, B)
        dout0 = self._bn_bwd("bnr0", out0, dr1, n, 32)
        self._lstm_bwd(0, X_tm, 6, 6, dout0, T, B, want_dx=False)
        # signal branch: d(sig) = dtot[:, 128:192]
        self.gemm(1, 0, L * Cc, 64, n, res, L * Cc, dtot, 192, self.g["sig_k"], 64, bo=128)
        self._call("nrvt_colsum", self._ptr(dtot, 128), n, 64, 192, self._ptr(self.g["sig_b"]), 0.0)
        dres = self.buf("dres", (n * L, Cc))
        self.gemm(0, 1, n, L * Cc, 64, dtot, 192, self.p["sig_k"], 64, dres, L * Cc, ao=128)
        self._call("nrvt_dropout", self._ptr(dres), C.c_void_p(mask.data_ptr()), n * L * Cc, keep_scale)
        dc2 = self._bn_bwd("bn2", c2, dres, n * L, Cc)
        db1 = self.buf("db1", (n * L, Cc))
        work = self.buf("bn_work", (4 * 256,), dtype=torch.float64)
        self._call("nrvt_conv1d_bwd", self._ptr(b1), self._ptr(self.p["conv2_k"]), self._ptr(c2), self._ptr(dc2), self._ptr(db1),
                   self._ptr(self.g["conv2_k"]), self._ptr(self.g["conv2_b"]), C.c_void_p(work.data_ptr()), n, L, Cc, Cc)
        dc1 = self._bn_bwd("bn1", c1, db1, n * L, Cc)
        self._call("nrvt_conv1d_bwd", self._ptr(S_tm), self._ptr(self.p["conv1_k"]), self._ptr(c1), self._ptr(dc1), None,
                   self._ptr(self.g["conv1_k"]), self._ptr(self.g["conv1_b"]), C.c_void_p(work.data_ptr()), n, L, 1, Cc)
        self._last_rows = {"bn1": n * L, "bn2": n * L, "bnr0": n, "bnr1": n, "bnr2": n}
        return stats

    @staticmethod
    def _metrics(stats, B):
        ce, hit, l2, l2hit = (float(v) for v in stats[:4].tolist())
        ce, l2 = ce / B, l2 / B
        # (metrics=['accuracy'] applies to both outputs in Keras: the l2_loss1 output gets binary_accuracy against its zero target)
        return {"loss": LOSS_WEIGHTS[0] * ce + LOSS_WEIGHTS[1] * l2, "final_out_loss": ce, "l2_loss1_loss": l2, "final_out_acc": hit / B,
                "l2_loss1_acc": l2hit / B}

    def _evaluate(self, S_tm, X_tm, y_d, B, T):
        """Validation pass: the predict graph (moving statistics, no dropout) through the same operators, and the two loss terms."""
        f = self._forward(S_tm, X_tm, B, T, False)
        stats = self.buf("stats", (4,), zero=True)
        probs = self.buf("probs", (B, self.n_class))
        dlog = self.buf("dlogits", (B, self.n_class))
        self._call("nrvt_softmax_ce", self._ptr(f["logits"]), C.c_void_p(y_d.data_ptr()), None, self._ptr(probs), self._ptr(dlog), self._ptr(stats),
                   B, self.n_class, 0.0)
        scratch_f, scratch_c = self.buf("ev_dfeat", (B, CENTER_DIM), zero=True), self.buf("ev_dcent", (self.n_class, CENTER_DIM), zero=True)
        self._call("nrvt_center_loss", self._ptr(f["feat"]), C.c_void_p(y_d.data_ptr()), self._ptr(self.p["centers"]), self._ptr(scratch_f),
                   self._ptr(scratch_c), self._ptr(stats), B, CENTER_DIM, 0.0)
        return self._metrics(stats, B)

    def predict(self, S, X):
        """att_model_predict.predict([S, X]) -> softmax probabilities [B, n_class] (the validation graph)."""
        B = np.asarray(X).shape[0]
        self.forward_backward(S, X, np.zeros(B), training=False)
        return self._buf["probs"].cpu().numpy()

    def _adam_step_size(self, t):
        return self.lr * math.sqrt(1.0 - self.beta2 ** t) / (1.0 - self.beta1 ** t)

    def _apply_device(self):
        """Adam on every trainable parameter + the BatchNormalization moving averages; step size from the device scalar"""
        for k, p in self.p.items():
            self._call("nrvt_adam", self._ptr(p), self._ptr(self.g[k]), self._ptr(self.m[k]), self._ptr(self.v[k]), p.numel(),
                       self._ptr(self._lr_t_d), self.beta1, self.beta2, self.eps)
        for name, rows in self._last_rows.items():
            Cn = self.s[name + "_mean"].numel()
            self._call("nrvt_ema", self._ptr(self.s[name + "_mean"]), self._ptr(self._buf[name + "_bm"]), Cn, BN_MOMENTUM, 1.0)
            self._call("nrvt_ema", self._ptr(self.s[name + "_var"]), self._ptr(self._buf[name + "_bv"]), Cn, BN_MOMENTUM,
                       rows / (rows - (1.0 + BN_EPS)))

    def apply_gradients(self):
        """One Adam step on every trainable parameter and the BatchNormalization moving-average updates of the last batch."""
        self.iterations += 1
        self._lr_t_d.fill_(self._adam_step_size(self.iterations))
        self._apply_device()

    def train_on_batch(self, S, X, y, class_weight=None, dropout_mask=None, graph=False):
        """One training step.  graph=True: the device work of a step (forward, backward, Adam: ~600 launches) is captured into a
        CUDA graph the third time a batch size is seen and replayed from then on; the inputs go into the graph's static tensors,
        the step number (dropout mask) and the Adam step size into device scalars.  Same arithmetic either way."""
        if not graph or dropout_mask is not None:
            out = self.forward_backward(S, X, y, class_weight, True, dropout_mask)
            self.apply_gradients()
            return out
        torch = self._torch
        S_tm, X_tm, y_d, B, T = self._upload(S, X, y)
        key = (B, T, self.class_weight_mode if class_weight is not None else None)
        st = self._graphs.setdefault(key, {"seen": 0})
        st["seen"] += 1
        self._step_d.fill_(self.iterations)
        self.iterations += 1
        self._lr_t_d.fill_(self._adam_step_size(self.iterations))
        if "graph" not in st:
            if st["seen"] < 3:              # warm-up: every buffer of this batch size gets allocated outside the capture
                stats = self._train_device(S_tm, X_tm, y_d, B, T, self._class_weight_tensor(class_weight), None)
                self._apply_device()
                return self._metrics(stats, B)
            st["in"] = (S_tm.clone(), X_tm.clone(), y_d.clone())
            st["cw"] = self._class_weight_tensor(class_weight)
            torch.cuda.synchronize(self.dev)
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    st["stats"] = self._train_device(st["in"][0], st["in"][1], st["in"][2], B, T, st["cw"], None)
                    self._apply_device()
            except Exception as e:          # capture refused (nothing has executed): keep launching this batch size eagerly, loudly
                print("[nanoreviser_b200.train] CUDA graph capture failed (%s); launching the steps one by one" % e)
                st["seen"] = -(1 << 60)
                stats = self._train_device(S_tm, X_tm, y_d, B, T, st["cw"], None)
                self._apply_device()
                return self._metrics(stats, B)
            st["graph"] = g                 # (the capture itself does not execute: the replay below is this batch's step)
        for dst, src in zip(st["in"], (S_tm, X_tm, y_d)):
            dst.copy_(src)
        st["graph"].replay()
        return self._metrics(st["stats"], B)

    # -- Model.fit -------------------------------------------------------------------------------------------------------------
    def fit(self, inputs, targets=None, class_weight=None, validation_split=0.0, shuffle=True, epochs=1, batch_size=32, verbose=1,
            seed=0):
        """fit([signal_x[..., None], x, y], [y, zeros], class_weight=..., validation_split=..., shuffle=True, epochs=..., batch_size=...)
        -> history dictionary (lists per epoch), Keras semantics: the LAST validation_split fraction of the samples is held out
        before shuffling; every epoch visits the remaining samples once in a fresh random order; metrics are sample-weighted means
        over the epoch's batches."""
        S, X, y = inputs[0], inputs[1], np.asarray(inputs[2]).reshape(-1)
        S, X = np.asarray(S), np.asarray(X)
        N = X.shape[0]
        n_train = int(N * (1.0 - float(validation_split))) if validation_split else N      # keras/engine/training.py: split_at
        n_val = N - n_train
        if n_train <= 0:
            raise ValueError("no training samples")
        rng = np.random.default_rng(seed)
        history: Dict[str, List[float]] = {}
        keys = ("loss", "final_out_loss", "l2_loss1_loss", "final_out_acc", "l2_loss1_acc")
        for ep in range(int(epochs)):
            t0 = time.time()
            order = rng.permutation(n_train) if shuffle else np.arange(n_train)
            acc = dict.fromkeys(keys, 0.0)
            for a in range(0, n_train, int(batch_size)):
                idx = np.sort(order[a:a + int(batch_size)])
                m = self.train_on_batch(S[idx], X[idx], y[idx], class_weight, graph=self.use_graph)
                for k in keys:
                    acc[k] += m[k] * len(idx)
            row = {k: acc[k] / n_train for k in keys}
            if n_val:
                vacc = dict.fromkeys(keys, 0.0)
                for a in range(n_train, N, int(batch_size)):
                    b = min(N, a + int(batch_size))
                    m = self.forward_backward(S[a:b], X[a:b], y[a:b], training=False)
                    for k in keys:
                        vacc[k] += m[k] * (b - a)
                row.update({"val_" + k: vacc[k] / n_val for k in keys})
            for k, v in row.items():
                history.setdefault(k, []).append(float(v))
            if verbose:
                print("Epoch %d/%d - %ds - " % (ep + 1, int(epochs), int(time.time() - t0)) +
                      " - ".join("%s: %.4f" % (k, v) for k, v in row.items()))
        return history

    # -- weights out -----------------------------------------------------------------------------------------------------------
    def get_weights(self) -> ModelWeights:
        a = lambda k: self.p[k].cpu().numpy().copy()
        bn = lambda n: np.stack([a(n + "_g"), a(n + "_b"), self.s[n + "_mean"].cpu().numpy(), self.s[n + "_var"].cpu().numpy()])
        lstm = [tuple(LstmDir(a("l%d%d_k" % (l, d)), a("l%d%d_r" % (l, d)), a("l%d%d_b" % (l, d))) for d in range(2)) for l in range(4)]
        return ModelWeights(window=self.window, n_class=self.n_class, conv1_k=a("conv1_k"), conv1_b=a("conv1_b"), bn1=bn("bn1"),
                            conv2_k=a("conv2_k"), conv2_b=a("conv2_b"), bn2=bn("bn2"), sig_dense_k=a("sig_k"), sig_dense_b=a("sig_b"),
                            lstm=lstm, bn_rnn=[bn("bnr0"), bn("bnr1"), bn("bnr2")], dense1_k=a("d1_k"), dense1_b=a("d1_b"),
                            dense2_k=a("d2_k"), dense2_b=a("d2_b"), main_k=a("m_k"), main_b=a("m_b"), feat_k=a("f_k"), feat_b=a("f_b"),
                            final_k=a("o_k"), final_b=a("o_b"))

    def save_weights(self, path: str):
        """att_model_predict.save_weights(path): Keras 2.2.4 layout (layer_names / weight_names attributes, /<layer>/<weight name>)."""
        save_predict_weights(self.get_weights(), path)


_RNN_NAMES = ("read_rnn1", "read_rnn11", "total_rnn1", "total_rnn2")


def save_predict_weights(w: ModelWeights, path: str):
    """The predict model's 24 layers in the order and with the names Keras gives them in lstmmodel.py:32-81 (the files under model/)."""
    S = lambda names: np.array([n.encode() for n in names], dtype="S%d" % max(len(n) for n in names)) if names else np.zeros(0, "S1")

    def layer(name, arrays):            # arrays: [(weight name relative to the layer, ndarray)]
        tree: dict = {}
        for wn, arr in arrays:
            node = tree.setdefault(name, {})
            parts = wn.split("/")
            for p in parts[:-1]:
                node = node.setdefault(p, {})
            node[parts[-1]] = np.ascontiguousarray(arr, np.float32)
        tree["__attrs__"] = {"weight_names": S(["%s/%s" % (name, wn) for wn, _ in arrays])}
        return tree

    kb = lambda k, b: [("kernel:0", k), ("bias:0", b)]
    bn = lambda a: [("gamma:0", a[0]), ("beta:0", a[1]), ("moving_mean:0", a[2]), ("moving_variance:0", a[3])]

    def bidir(l):
        out = []
        for d, prefix in enumerate(("forward_", "backward_")):
            ld = w.lstm[l][d]
            out += [("%s%s/kernel:0" % (prefix, _RNN_NAMES[l]), ld.kernel), ("%s%s/recurrent_kernel:0" % (prefix, _RNN_NAMES[l]), ld.recurrent),
                    ("%s%s/bias:0" % (prefix, _RNN_NAMES[l]), ld.bias)]
        return out

    layers = [("signal_input", []), ("time_distributed_1", kb(w.conv1_k, w.conv1_b)), ("time_distributed_2", bn(w.bn1)),
              ("time_distributed_3", kb(w.conv2_k, w.conv2_b)), ("read_input", []), ("time_distributed_4", bn(w.bn2)),
              ("bidirectional_1", bidir(0)), ("add_1", []), ("batch_normalization_3", bn(w.bn_rnn[0])), ("dropout_1", []),
              ("bidirectional_2", bidir(1)), ("time_distributed_5", []), ("batch_normalization_4", bn(w.bn_rnn[1])),
              ("time_distributed_6", kb(w.sig_dense_k, w.sig_dense_b)), ("concatenate_1", []), ("bidirectional_3", bidir(2)),
              ("batch_normalization_5", bn(w.bn_rnn[2])), ("bidirectional_4", bidir(3)), ("dense_1", kb(w.dense1_k, w.dense1_b)),
              ("dense_2", kb(w.dense2_k, w.dense2_b)), ("main_out", kb(w.main_k, w.main_b)), ("flatten_2", []),
              ("feature", kb(w.feat_k, w.feat_b)), ("final_out", kb(w.final_k, w.final_b))]
    tree: dict = {name: layer(name, arrays) for name, arrays in layers}
    tree["__attrs__"] = {"layer_names": S([n for n, _ in layers]), "backend": "tensorflow", "keras_version": "2.2.4"}
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as fp:
        fp.write(h5write.write_tree(tree))
