"""CPU ORACLE for the NanoReviser revision-inference path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  The product (``nanoreviser_b200``) never does: it fails loudly when
the CUDA library is missing.

What this restates (all citations relative to ``/root/reference``):

* A1  ``get_read_data``           nanorevutils/nanorev_fast5_handeler.py:39-150   -> :func:`get_read_data`
* A2  ``signal_segmentation``     nanorevutils/preprocessing.py:85-170            -> :func:`signal_segmentation`
* A3  feature columns             NanoReviser.py:120-125, nanorevtrainutils.py:159-169 -> :func:`feature_columns`
* A4  sliding windows             nanorevtrainutils.py:197-213                    -> :func:`make_windows`
* A5-A8 Keras-2.2.4 graph         nanorevcnn.py:17-38, lstmmodel.py:32-133        -> :func:`forward`
* A9-A10 label maps, two-model merge  output_handeler.py:83,104-122               -> :func:`get_base_1`
* A11 writers                     output_handeler.py:26-62, NanoReviser.py:137,163 -> :func:`prep_read_fasta` ...

PINNING STATUS (see oracle/pin_against_reference.py, which runs in the build container where
``/root/reference`` is mounted):

* A2, ``get_base_color`` / ``get_base_label``: checked against the reference module itself,
  imported unchanged (``nanorevutils.preprocessing`` is numpy-only) on all 105 fixture reads.
* A10, A11 and ``label_to_base``: checked against the reference functions AST-extracted from
  ``output_handeler.py`` (the module cannot be imported: ``import keras``).
* A1: checked against the reference ``get_read_data`` source executed unchanged with ``h5py`` /
  ``albacore`` shimmed by ``nanoreviser_b200.h5mini`` (AST-extracted), and against the in-file
  Albacore ``Fastq`` known answer ``bases == Fastq_seq[2:-2]``.
* A5-A8 (the Keras graph): **PARITY UNPINNED**.  Keras 2.2.4 / TensorFlow 1.12 / h5py cannot be
  installed here (no network) and the reference's own tests hold no numeric vectors
  (unitest/check_nanoreviser.py:23-41 greps a log for "Congratulations").  The restatement follows
  Keras-2.2.4 semantics (LSTM gate order i,f,c,o; recurrent_activation hard_sigmoid; BN eps 1e-3;
  Conv1D cross-correlation 'same'; Add broadcasts the 1-channel input) and is sanity-checked by
  (i) the trained weights reproducing the basecalled base on 96-99.7 % of positions, (ii) the
  survey's independent probe values, (iii) fp32-vs-fp64 agreement.

Composition of the NN path (the shipped CLI never runs it -- SURVEY.md F1; every item is a builder
decision, recorded here verbatim):

D1 ``W = feature.kernel.shape[0] // 6`` (= 11; the files are named win13).
D2 windows ``i = 0 .. N-W-1`` exactly as nanorevtrainutils.py:198 (M = N-W predictions, window i is
   centred on base ``i + 5``).
D3 ``y1 = argmax(P1)``, ``y2 = argmax(P2)``, first max on ties (numpy semantics).
D4 ``core = get_base_1(bases[5:5+M], y1, y2 + 2)`` -- the reference function verbatim; model-2 class k
   is label k+1 (nanorevtrainutils.py:213) and get_base_1 subtracts 1 itself (output_handeler.py:106).
   ``revised = bases[:5] + core + bases[5+M:]``; reads with ``N <= W`` pass through unchanged.
D5 any per-read failure -> the original ``event_bases`` are written (NanoReviser.py:146-154).
D6 ``-F fastq`` qualities have no NN-path definition in the reference; parity is claimed for fasta.
D7 ``-S`` defaults to 'human'; model paths per NanoReviser.py:192-193.
D8 ``scale == 0`` (MAD of the signal is zero; the reference would divide by zero and carry NaNs
   silently) is a per-read failure -> D5.
"""
from __future__ import annotations

import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from nanoreviser_b200 import h5mini  # noqa: E402  (I/O only: HDF5 container parsing)

SIGNAL_LEN = 50
SET_BEF = 5

# --------------------------------------------------------------------------------------
# A9: label / colour maps  (preprocessing.py:173-180, output_handeler.py:83)
# --------------------------------------------------------------------------------------
label_to_base = {5: 'A', 4: 'G', 3: 'T', 2: 'C', 1: '-', 0: 'D'}


def get_base_color(base):
    return {'A': 250, 'G': 180, 'T': 100, 'C': 30}.get(base, 0)


def get_base_label(base):
    return {'A': 5, 'G': 4, 'T': 3, 'C': 2, '-': 1, 'D': 0}.get(base, 0)


# --------------------------------------------------------------------------------------
# A1: events -> bases   (nanorev_fast5_handeler.py:39-150)
# --------------------------------------------------------------------------------------
def _loose_version(v):
    """distutils.version.LooseVersion(v).version (the reference's comparison, fast5_handeler.py:65-68): the string is cut into
    runs of digits / runs of letters, dots dropped, digit runs become ints; compared as lists."""
    import re
    if isinstance(v, (bytes, np.bytes_)):
        v = bytes(v).decode()
    parts = [x for x in re.split(r'(\d+|[a-z]+|\.)', str(v)) if x and x != '.']
    return [int(x) if x.isdigit() else x for x in parts]


def get_read_data(fast5_fn, basecall_group='Basecall_1D_000', basecall_subgroup='BaseCalled_template'):
    f = h5mini.File(fast5_fn, 'r')
    grp = f['/Analyses/' + basecall_group]
    version = grp.attrs['version'] if 'version' in grp.attrs else b'0.0'
    called = f['/Analyses/' + basecall_group + '/' + basecall_subgroup + '/Events'][()]
    if _loose_version(version) <= _loose_version('0.0'):   # :65-73 legacy tables (LooseVersion(v) <= LooseVersion('0.0'))
        raw_attrs = dict(list(f['/Raw/Reads/'].values())[0].attrs.items())
        called['start'] = called['start'] * 4000 - raw_attrs['start_time']
        called['length'] = called['length'] * 4000
    start, bases, ab_mean, ab_std = [], [], [], []
    n = len(called)
    for k in range(n - 1, -1, -1):                     # :84-114, reversed iteration
        model_state = called['model_state'][k].decode()
        start_l = int(called['start'][k])
        move = int(called['move'][k])
        p_mean, p_std = called['mean'][k], called['stdv'][k]
        if move == 0:
            continue
        if move == 2:
            start.append(start_l + 2); bases.append(model_state[2]); ab_mean.append(p_mean); ab_std.append(p_std)
            start.append(start_l); bases.append(model_state[1]); ab_mean.append(p_mean); ab_std.append(p_std)
        else:
            start.append(start_l); bases.append(model_state[2]); ab_mean.append(p_mean); ab_std.append(p_std)
    start, bases, ab_mean, ab_std = start[::-1], bases[::-1], ab_mean[::-1], ab_std[::-1]
    length = list(np.diff(start))                      # :121-126
    length.append(3. if start[-1] - start[-2] < 5 else 5.)
    read_name = list(f['/Raw/Reads/'].items())[0][0]
    signal = f['/Raw/Reads/' + read_name + '/Signal'][()]
    f.close()
    if len(signal) < int(start[-1] + length[-1]):      # :142-143
        raise RuntimeError('Signal is shorter than the Events')
    a0 = start[0]
    return a0, np.array(start) - a0, np.array(length), bases, signal, ab_mean, ab_std


def read_fastq_dataset(fast5_fn, basecall_group='Basecall_1D_000', basecall_subgroup='BaseCalled_template'):
    f = h5mini.File(fast5_fn, 'r')
    fq = bytes(f['/Analyses/' + basecall_group + '/' + basecall_subgroup + '/Fastq'][()]).decode('utf8')
    f.close()
    return fq.split('\n')


# --------------------------------------------------------------------------------------
# A2: signal_segmentation   (preprocessing.py:85-170), float64 exactly as the reference
# --------------------------------------------------------------------------------------
def signal_segmentation(raw_signal, starts, last_dur, query_len=50):
    raw = np.asarray(raw_signal)
    sig = raw.astype(np.float64)
    S = len(sig)
    half = query_len // 2
    shift = np.median(sig)                              # :95-97  (no 1.4826 factor, no epsilon)
    scale = np.median(np.abs(sig - shift))
    starts = np.asarray(starts).astype(np.int64)
    N = len(starts)
    ends = np.empty(N, dtype=np.int64)
    ends[:-1] = starts[1:]
    ends[-1] = starts[-1] + int(last_dur)               # :103-104,134
    win = np.zeros((N, query_len), dtype=np.float64)
    mean = np.empty(N, dtype=np.float64)
    std = np.empty(N, dtype=np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        norm = (sig - shift) / scale                    # :119 elementwise, same fp64 ops
    for j in range(N):
        st = int(starts[j])
        lo = 0 if st - half <= 0 else st - half         # :111-118
        hi = S if st + half >= S else st + half
        seg = norm[lo:hi]
        pad = query_len - len(seg)
        if pad > 0:
            left = pad // 2 + 1 if pad % 2 else pad // 2   # :120-131  odd pad: (p//2+1, p//2)
            win[j, left:left + len(seg)] = seg
        else:
            win[j, :] = seg
        tmp = sig[st:int(ends[j])]                      # raw, un-normalised :109,163
        with np.errstate(all='ignore'):
            mean[j] = np.mean(tmp)
            std[j] = np.std(tmp)
    return win, mean, std, shift, scale


# --------------------------------------------------------------------------------------
# A3: feature columns   (NanoReviser.py:124-125; nanorevtrainutils.py:159-169)
# --------------------------------------------------------------------------------------
def feature_columns(bases, seg_mean, seg_std, shift, scale, length, ev_mean, ev_std):
    colour = np.array([get_base_color(b) for b in bases], dtype=np.float64) / 300.0
    with np.errstate(divide='ignore', invalid='ignore'):
        x = np.vstack([colour,
                       np.asarray(seg_mean, dtype=np.float64) / shift,
                       np.asarray(seg_std, dtype=np.float64) / scale,
                       np.asarray(length, dtype=np.float64) / 10.0,
                       np.asarray(ev_mean, dtype=np.float64),
                       np.asarray(ev_std, dtype=np.float64)])
    return np.array(x.T, dtype='float')                 # [N, 6] float64


# --------------------------------------------------------------------------------------
# A4: sliding windows   (nanorevtrainutils.py:197-209)
# --------------------------------------------------------------------------------------
def make_windows(x, signal_x, W):
    n = len(x) - W
    if n <= 0:
        return np.zeros((0, W, x.shape[1])), np.zeros((0, W, signal_x.shape[1]))
    idx = np.arange(n)[:, None] + np.arange(W)[None, :]
    return x[idx], signal_x[idx]


# --------------------------------------------------------------------------------------
# A5-A8: the Keras-2.2.4 graph   (nanorevcnn.py:17-38, lstmmodel.py:32-133)
# --------------------------------------------------------------------------------------
BN_EPS = 1e-3   # keras.layers.BatchNormalization default epsilon


def _bn(x, bn, dt):
    gamma, beta, mean, var = (bn[i].astype(dt) for i in range(4))
    inv = gamma / np.sqrt(var + dt(BN_EPS))            # tf.nn.batch_normalization form
    return x * inv + (beta - mean * inv)


def _hard_sigmoid(x):
    return np.clip(0.2 * x + 0.5, 0.0, 1.0)            # keras.backend.tensorflow_backend.hard_sigmoid


def _conv1d_same_relu(x, k, b):
    """x [n, L, cin], k [3, cin, cout] (Keras layout; cross-correlation, zero pad 1/1)."""
    n, L, cin = x.shape
    xp = np.zeros((n, L + 2, cin), dtype=x.dtype)
    xp[:, 1:-1] = x
    y = xp[:, 0:L] @ k[0] + xp[:, 1:L + 1] @ k[1] + xp[:, 2:L + 2] @ k[2] + b
    return np.maximum(y, 0)


def cnn_branch(m, sig, dt=np.float32):
    """identity_Block + TD(Flatten) + TD(Dense 64) applied to a set of 50-sample signals.

    sig [n, 50] -> [n, 64].  (nanorevcnn.py:29-38, lstmmodel.py:35-41)
    """
    x0 = sig.astype(dt)[:, :, None]
    x = _conv1d_same_relu(x0, m.conv1_k.astype(dt), m.conv1_b.astype(dt))
    x = _bn(x, m.bn1, dt)
    x = _conv1d_same_relu(x, m.conv2_k.astype(dt), m.conv2_b.astype(dt))
    x = _bn(x, m.bn2, dt)
    x = x + x0                                          # Add broadcasts the 1-channel input
    flat = x.reshape(x.shape[0], -1)                    # index = pos*8 + ch
    return flat @ m.sig_dense_k.astype(dt) + m.sig_dense_b.astype(dt)


def _lstm_dir(x, d, reverse, dt):
    """x [B, T, in] -> h sequence [B, T, u] in input time order."""
    B, T, _ = x.shape
    u = d.recurrent.shape[0]
    Wk, Wr, b = d.kernel.astype(dt), d.recurrent.astype(dt), d.bias.astype(dt)
    zin = x @ Wk + b
    h = np.zeros((B, u), dtype=dt)
    c = np.zeros((B, u), dtype=dt)
    out = np.empty((B, T, u), dtype=dt)
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        z = zin[:, t] + h @ Wr
        i = _hard_sigmoid(z[:, :u])
        f = _hard_sigmoid(z[:, u:2 * u])
        g = np.tanh(z[:, 2 * u:3 * u])
        o = _hard_sigmoid(z[:, 3 * u:])
        c = f * c + i * g
        h = o * np.tanh(c)
        out[:, t] = h
    return out


def _bilstm(x, pair, dt):
    return np.concatenate([_lstm_dir(x, pair[0], False, dt), _lstm_dir(x, pair[1], True, dt)], axis=-1)


def forward_windows(m, S, X, dt=np.float32, return_logits=False):
    """``model.predict([S[..., None], X])`` with per-window CNN recompute (what Keras does).

    S [B, W, 50], X [B, W, 6] -> softmax probabilities [B, n_class].
    """
    B, W, _ = X.shape
    sig_feat = cnn_branch(m, S.reshape(B * W, SIGNAL_LEN), dt).reshape(B, W, 64)
    return _trunk(m, sig_feat, X.astype(dt), dt, return_logits)


def _trunk(m, sig_feat, X, dt, return_logits=False):
    r1 = _bn(_bilstm(X, m.lstm[0], dt), m.bn_rnn[0], dt)
    r2 = _bn(_bilstm(r1, m.lstm[1], dt), m.bn_rnn[1], dt)
    tot = np.concatenate([r2, sig_feat], axis=-1)       # order [read_rnn2(128), signal(64)]
    t1 = _bn(_bilstm(tot, m.lstm[2], dt), m.bn_rnn[2], dt)
    t2 = _bilstm(t1, m.lstm[3], dt)
    d = np.maximum(t2 @ m.dense1_k.astype(dt) + m.dense1_b.astype(dt), 0)
    d = np.maximum(d @ m.dense2_k.astype(dt) + m.dense2_b.astype(dt), 0)
    d = np.maximum(d @ m.main_k.astype(dt) + m.main_b.astype(dt), 0)
    flat = d.reshape(d.shape[0], -1)                    # index = t*6 + k
    feat = np.maximum(flat @ m.feat_k.astype(dt) + m.feat_b.astype(dt), 0)
    logits = feat @ m.final_k.astype(dt) + m.final_b.astype(dt)
    if return_logits:
        return logits
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


def forward_read(m, sig_win, x, dt=np.float32, chunk=4096, return_logits=False):
    """Batched oracle: CNN once per base, then all N-W windows of one read (same maths)."""
    W = m.window
    N = len(x)
    M = N - W
    nc = m.n_class
    if M <= 0:
        return np.zeros((0, nc), dtype=dt)
    sig_feat = np.concatenate([cnn_branch(m, sig_win[a:a + 8192], dt) for a in range(0, N, 8192)])
    xs = x.astype(dt)
    out = np.empty((M, nc), dtype=dt)
    for a in range(0, M, chunk):
        b = min(M, a + chunk)
        idx = np.arange(a, b)[:, None] + np.arange(W)[None, :]
        out[a:b] = _trunk(m, sig_feat[idx], xs[idx], dt, return_logits)
    return out


# --------------------------------------------------------------------------------------
# A10: two-model merge   (output_handeler.py:104-122)
# --------------------------------------------------------------------------------------
def get_base_1(event_bases, y_pre, y_pre2):
    result = [label_to_base[int(y_pre[0])]]             # :107 leading symbol (may be the letter 'D')
    y_pre2 = np.asarray(y_pre2) - 1                     # :106
    for y_tmp, y_tmp2, base in zip(y_pre, y_pre2, event_bases):
        y_tmp = label_to_base.get(int(y_tmp), 0)
        y_tmp2 = label_to_base.get(int(y_tmp2), 0)
        if y_tmp == y_tmp2 and y_tmp in ['A', 'T', 'C', 'G']:
            result.append(y_tmp)
        elif y_tmp == 'D' and y_tmp2 in ['A', 'T', 'C', 'G']:
            result.append(base)
            result.append(y_tmp2)
        elif y_tmp == '-' and y_tmp2 == '-':
            continue
        else:
            result.append(base)
    return ''.join([t for t in result if t != '-'])      # :121


# --------------------------------------------------------------------------------------
# D6': fastq qualities of the NN path.  The reference has NO definition
# for them (-F fastq only exists on the Guppy path, output_handeler.py:86-102, and in the per-read
# fallback that copies the basecaller's Fastq[7:-7], NanoReviser.py:172-181); this is the builder's
# definition, restated here so that the CUDA path can be checked byte for byte:
#   * Phred of a window and model: Q(p) = #{k in 1..60 : float32(1) - p <= float32(10**(-k/10))} with
#     p = the softmax value of the argmax class as float32 (i.e. floor(-10 log10(1-p)) capped at 60,
#     evaluated with a threshold table so that it is exact and portable);
#   * every symbol get_base_1 emits BECAUSE of the models (agreed base, inserted base) gets
#     min(Q1, Q2) of its window, the leading label symbol gets Q1 of window 0;
#   * every base that passes through (edges, disagreement, kept base before an insertion, failed
#     read) keeps the basecaller's quality when it is known, else Phred 40 (the former constant 'I').
# Output characters are Phred + 33.
# --------------------------------------------------------------------------------------
PHRED_MAX = 60
PHRED_PASS = 40
PHRED_THRESH = np.array([10.0 ** (-k / 10.0) for k in range(1, PHRED_MAX + 1)]).astype(np.float32)


def phred_of_prob(p):
    """p: float32 array of argmax-class probabilities -> uint8 Phred 0..60."""
    e = np.float32(1.0) - np.asarray(p, dtype=np.float32)
    return (e[..., None] <= PHRED_THRESH).sum(axis=-1).astype(np.uint8)


def get_qual_1(event_bases, y_pre, y_pre2, q1, q2, pass_q):
    """Phred values parallel to get_base_1(event_bases, y_pre, y_pre2): same branches, same order
    (output_handeler.py:104-122), one value per emitted symbol, dropped where get_base_1 drops '-'."""
    res = []
    lead = label_to_base[int(y_pre[0])]
    if lead != '-':
        res.append(int(q1[0]))
    y_pre2 = np.asarray(y_pre2) - 1
    for k, (y_tmp, y_tmp2, base) in enumerate(zip(y_pre, y_pre2, event_bases)):
        y_tmp = label_to_base.get(int(y_tmp), 0)
        y_tmp2 = label_to_base.get(int(y_tmp2), 0)
        qm = min(int(q1[k]), int(q2[k]))
        if y_tmp == y_tmp2 and y_tmp in ['A', 'T', 'C', 'G']:
            res.append(qm)
        elif y_tmp == 'D' and y_tmp2 in ['A', 'T', 'C', 'G']:
            if base != '-':
                res.append(int(pass_q[k]))
            res.append(qm)
        elif y_tmp == '-' and y_tmp2 == '-':
            continue
        else:
            if base != '-':
                res.append(int(pass_q[k]))
    return res


def revise_quality(bases, y1, y2, p1max, p2max, window, qual_in=None, revised_ok=True):
    """Quality string (Phred+33) of the whole revised read under composition D4; p1max / p2max are the
    float32 argmax probabilities of the N-W windows; qual_in the basecaller's Phred values per base or None."""
    N = len(bases)
    pass_q = np.full(N, PHRED_PASS, dtype=np.int64) if qual_in is None else np.minimum(np.asarray(qual_in, dtype=np.int64), 93)
    M = N - window
    if not revised_ok or M <= 0:
        return ''.join(chr(int(q) + 33) for q in pass_q)
    q1, q2 = phred_of_prob(p1max), phred_of_prob(p2max)
    core = get_qual_1(bases[SET_BEF:SET_BEF + M], y1, np.asarray(y2) + 2, q1, q2, pass_q[SET_BEF:SET_BEF + M])
    allq = list(pass_q[:SET_BEF]) + core + list(pass_q[SET_BEF + M:])
    return ''.join(chr(int(q) + 33) for q in allq)


# --------------------------------------------------------------------------------------
# A11: writers   (output_handeler.py:26-62; file names NanoReviser.py:137,163)
# --------------------------------------------------------------------------------------
def fasta_text(fast5_fn, bases):
    return ">" + fast5_fn.split('/')[-1].replace(' ', '|||') + '\n' + ''.join(bases)


def fastq_text(fast5_fn, bases, qul):
    return "@" + fast5_fn.split('/')[-1].replace(' ', '|||') + '\n' + ''.join(bases) + '+\n' + ''.join(qul)


def out_filename(output_dir, fast5_fn_sg, fmt):
    return output_dir + fast5_fn_sg.split('.')[0] + '_out.' + fmt


# --------------------------------------------------------------------------------------
# composition D1-D8
# --------------------------------------------------------------------------------------
def revise_arrays(m1, m2, bases, starts, length, signal_from_a0, ev_mean, ev_std, dt=np.float32,
                  want=()):
    """Full NN path for one read, from the outputs of get_read_data.  Returns a dict."""
    W = m1.window
    N = len(bases)
    res = {'status': 0}
    bases = list(bases)
    win, seg_mean, seg_std, shift, scale = signal_segmentation(signal_from_a0, starts, int(length[-1]))
    res.update(shift=shift, scale=scale)
    if N <= W:
        res.update(revised=''.join(bases), status=1)
        return res
    if not (scale > 0) or not np.isfinite(shift):
        res.update(revised=''.join(bases), status=2)      # D8
        return res
    x = feature_columns(bases, seg_mean, seg_std, shift, scale, length, ev_mean, ev_std)
    P1 = forward_read(m1, win, x, dt)
    P2 = forward_read(m2, win, x, dt)
    y1 = np.argmax(P1, axis=1)
    y2 = np.argmax(P2, axis=1)
    M = N - W
    core = get_base_1(bases[SET_BEF:SET_BEF + M], y1, y2 + 2)
    res['revised'] = ''.join(bases[:SET_BEF]) + core + ''.join(bases[SET_BEF + M:])
    res.update(y1=y1, y2=y2)
    if 'probs' in want:
        res.update(P1=P1, P2=P2)
    if 'features' in want:
        res.update(x=x, sig_win=win, seg_mean=seg_mean, seg_std=seg_std)
    return res


def revise_fast5(m1, m2, fast5_fn, dt=np.float32, want=()):
    a0, starts, length, bases, signal, ev_mean, ev_std = get_read_data(fast5_fn)
    res = revise_arrays(m1, m2, bases, starts, length, signal[int(a0):], ev_mean, ev_std, dt, want)
    res.update(a0=a0, starts=starts, length=length, bases=bases, ev_mean=ev_mean, ev_std=ev_std)
    return res
