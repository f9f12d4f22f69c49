"""INDEPENDENT second implementation of rows A5-A8 (test infrastructure; not collected as a test module).

Purpose ("break the common mode"): the CUDA path and the CPU oracle both obtain their weights through
``nanoreviser_b200.h5mini`` + ``nanoreviser_b200.weights`` (layers mapped by POSITION in the ``layer_names`` attribute) and the
oracle's forward is ``oracle/nanorev_oracle.py``.  A mis-parsed dataset, a fwd/bwd swap or a layer mapped to the wrong role would be
common to both and every GPU-vs-oracle test would still pass.  This module shares NOTHING with them:

* :class:`H5Scan` -- its own reader of the HDF5 subset Keras 2.2.4 ``save_weights`` files use (superblock v0, v1 object headers,
  symbol-table groups, contiguous little-endian float32 datasets), written with ``struct`` against the HDF5 file-format
  specification; it ignores attributes entirely and enumerates datasets by walking the group tree.
* :func:`load_by_name` -- maps datasets to roles by their Keras PATH NAMES (``.../forward_total_rnn1/recurrent_kernel:0``) and
  SHAPES, never by position.
* :class:`KerasGraph` -- the graph of ``nanorevutils/lstmmodel.py:32-133`` + ``nanorevcnn.py:17-38`` in torch float64 with
  ``torch.nn.functional`` primitives (conv1d, batch_norm, linear), each Keras-2.2.4 convention being a named, switchable option so
  that the survey's convention ablation can be re-run (tests/test_indep_forward.py).
"""
import re
import struct

import numpy as np
import torch
import torch.nn.functional as F

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Scan:
    """walk() -> {"/group/.../dataset": float32 ndarray} for every contiguous float32 dataset of the file"""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        if self.buf[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file")
        ver, = struct.unpack_from("<B", self.buf, 8)
        so, sl = struct.unpack_from("<BB", self.buf, 13)
        if ver != 0 or so != 8 or sl != 8:
            raise ValueError("superblock version %d / offset size %d not handled" % (ver, so))
        # superblock v0: 24 bytes of fixed fields, 4 addresses, then the root group's symbol-table entry
        self.root_header, = struct.unpack_from("<Q", self.buf, 24 + 4 * 8 + 8)

    # -- object headers ------------------------------------------------------------------------------------------------
    def messages(self, addr):
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", self.buf, addr)
        if ver != 1:
            raise ValueError("object header version %d" % ver)
        spans = [(addr + 16, hsize)]
        out = []
        while spans and len(out) < nmsg:
            pos, size = spans.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", self.buf, pos)
                body = pos + 8
                if mtype == 0x10:                                  # continuation
                    spans.append(struct.unpack_from("<QQ", self.buf, body))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    # -- groups --------------------------------------------------------------------------------------------------------
    def _heap_string(self, heap_addr, off):
        if self.buf[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("local heap signature")
        seg, = struct.unpack_from("<Q", self.buf, heap_addr + 24)
        s = seg + off
        return self.buf[s:self.buf.index(b"\x00", s)].decode()

    def _tree_entries(self, node, heap):
        sig = self.buf[node:node + 4]
        if sig == b"SNOD":
            n, = struct.unpack_from("<H", self.buf, node + 6)
            for i in range(n):
                name_off, obj = struct.unpack_from("<QQ", self.buf, node + 8 + 40 * i)
                yield self._heap_string(heap, name_off), obj
        elif sig == b"TREE":
            ntype, _level, used = struct.unpack_from("<BBH", self.buf, node + 4)
            if ntype != 0:
                raise ValueError("not a group B-tree")
            for i in range(used):
                child, = struct.unpack_from("<Q", self.buf, node + 24 + 16 * i + 8)
                yield from self._tree_entries(child, heap)
        else:
            raise ValueError("bad group node signature %r" % sig)

    def children(self, header_addr):
        for mtype, body, _ in self.messages(header_addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", self.buf, body)
                return list(self._tree_entries(btree, heap))
        return None                                                    # not a group

    # -- datasets ------------------------------------------------------------------------------------------------------
    def dataset(self, header_addr):
        shape = dtype_ok = addr = None
        for mtype, body, _ in self.messages(header_addr):
            if mtype == 0x01:
                ver, rank = struct.unpack_from("<BB", self.buf, body)
                dims_at = body + (8 if ver == 1 else 4)
                shape = struct.unpack_from("<%dQ" % rank, self.buf, dims_at)
            elif mtype == 0x03:
                cls = self.buf[body] & 0x0F
                size, = struct.unpack_from("<I", self.buf, body + 4)
                byte_order_be = self.buf[body + 1] & 1
                dtype_ok = (cls == 1 and size == 4 and not byte_order_be)
            elif mtype == 0x08:
                ver, lclass = struct.unpack_from("<BB", self.buf, body)
                if ver == 3 and lclass == 1:
                    addr, = struct.unpack_from("<Q", self.buf, body + 2)
        if shape is None or not dtype_ok or addr is None or addr == UNDEF:
            return None
        n = int(np.prod(shape)) if shape else 1
        return np.frombuffer(self.buf, dtype="<f4", count=n, offset=addr).reshape(shape).copy()

    def walk(self):
        out = {}

        def rec(addr, path, depth):
            if depth > 8:
                raise ValueError("group tree too deep")
            kids = self.children(addr)
            if kids is None:
                d = self.dataset(addr)
                if d is not None:
                    out[path] = d
                return
            for name, child in kids:
                rec(child, path + "/" + name, depth + 1)

        rec(self.root_header, "", 0)
        return out


def _suffix(name):
    m = re.search(r"_(\d+)$", name)
    return int(m.group(1)) if m else 0


def load_by_name(path):
    """-> dict of float64 torch tensors keyed by ROLE, found by dataset path names and shapes only."""
    ds = H5Scan(path).walk()
    t = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64))
    W = {}
    # LSTMs: the Keras layer names are spelled out in the paths
    for layer in ("read_rnn1", "read_rnn11", "total_rnn1", "total_rnn2"):
        for direction in ("forward", "backward"):
            hits = {k: v for k, v in ds.items() if k.split("/")[-2] == "%s_%s" % (direction, layer)}
            if len(hits) != 3:
                raise ValueError("%s_%s: %d datasets" % (direction, layer, len(hits)))
            for k, v in hits.items():
                W["%s/%s/%s" % (layer, direction, k.split("/")[-1].split(":")[0])] = t(v)
    # everything else: by the leaf name and the shape
    by_layer = {}
    for k, v in ds.items():
        parts = k.strip("/").split("/")
        if "rnn" in parts[-2]:
            continue
        by_layer.setdefault(parts[-2], {})[parts[-1].split(":")[0]] = v
    convs = sorted((n for n, d in by_layer.items() if "kernel" in d and d["kernel"].ndim == 3), key=_suffix)
    assert [by_layer[n]["kernel"].shape for n in convs] == [(3, 1, 8), (3, 8, 8)], convs
    bns8 = sorted((n for n, d in by_layer.items() if "gamma" in d and d["gamma"].shape == (8,)), key=_suffix)
    assert len(bns8) == 2 and _suffix(convs[0]) < _suffix(bns8[0]) < _suffix(convs[1]) < _suffix(bns8[1])
    for role, n in (("conv1", convs[0]), ("conv2", convs[1])):
        W[role + "/kernel"], W[role + "/bias"] = t(by_layer[n]["kernel"]), t(by_layer[n]["bias"])
    for role, n in (("bn_conv1", bns8[0]), ("bn_conv2", bns8[1])):
        for p in ("gamma", "beta", "moving_mean", "moving_variance"):
            W[role + "/" + p] = t(by_layer[n][p])
    for width, role in ((32, "bn_read_rnn1"), (128, "bn_read_rnn11"), (256, "bn_total_rnn1")):
        hits = [n for n, d in by_layer.items() if "gamma" in d and d["gamma"].shape == (width,)]
        assert len(hits) == 1, (width, hits)
        for p in ("gamma", "beta", "moving_mean", "moving_variance"):
            W[role + "/" + p] = t(by_layer[hits[0]][p])
    dense_shapes = {(400, 64): "sig_dense", (128, 128): "dense1", (128, 32): "dense2", (32, 6): "main_out", (66, 16): "feature"}
    for n, d in by_layer.items():
        if "kernel" in d and d["kernel"].ndim == 2:
            shp = d["kernel"].shape
            role = dense_shapes.get(shp)
            if role is None and shp[0] == 16 and n.startswith("final_out"):
                role = "final_out"
            assert role is not None, (n, shp)
            assert role + "/kernel" not in W, role
            W[role + "/kernel"], W[role + "/bias"] = t(d["kernel"]), t(d["bias"])
    need = ["sig_dense", "dense1", "dense2", "main_out", "feature", "final_out"]
    assert all(r + "/kernel" in W for r in need)
    assert len(ds) == sum(1 for _ in W), "unmapped datasets: %d in file, %d mapped" % (len(ds), len(W))
    return W


DEFAULT_CONVENTIONS = dict(
    residual_add=True,            # nanorevcnn.py:37  Add()([x, input]) broadcasts the 1-channel input
    flatten="pos_major",          # TimeDistributed(Flatten) of [50, 8]: index = pos*8 + ch
    conv_flip=False,              # Keras Conv1D is a cross-correlation
    bn_before_relu=False,         # nanorevcnn.py:24-25: activation inside Conv1D, then BatchNormalization
    bn_eps=1e-3,                  # Keras default
    concat="read_then_signal",    # lstmmodel.py:48  concatenate([read_rnn2, signal_x_out])
    final_flatten="t_major",      # Flatten of [W, 6]: index = t*6 + k
    recurrent_activation="hard_sigmoid",   # Keras 2.2.4 LSTM default
    gate_order="ifco",            # Keras
    zero_signal=False,
)


class KerasGraph:
    def __init__(self, weights, **conventions):
        self.W = weights
        self.c = dict(DEFAULT_CONVENTIONS)
        unknown = set(conventions) - set(self.c)
        assert not unknown, unknown
        self.c.update(conventions)

    def _bn(self, x, role):
        W = self.W
        # F.batch_norm wants channels in dim 1
        xm = x.movedim(-1, 1)
        y = F.batch_norm(xm, W[role + "/moving_mean"], W[role + "/moving_variance"], W[role + "/gamma"], W[role + "/beta"],
                         training=False, eps=self.c["bn_eps"])
        return y.movedim(1, -1)

    def _conv(self, x, role):
        """x [n, L, cin] -> [n, L, cout], 'same' zero padding, relu (and BN in the chosen order)"""
        k = self.W[role + "/kernel"]                      # Keras (width, in, out)
        if self.c["conv_flip"]:
            k = k.flip(0)
        y = F.conv1d(x.transpose(1, 2), k.permute(2, 1, 0), self.W[role + "/bias"], padding=1).transpose(1, 2)
        bn = "bn_" + role
        if self.c["bn_before_relu"]:
            return torch.relu(self._bn(y, bn))
        return self._bn(torch.relu(y), bn)

    def signal_features(self, sig):
        """sig [n, 50] -> [n, 64]"""
        x0 = sig.unsqueeze(-1)
        x = self._conv(self._conv(x0, "conv1"), "conv2")
        if self.c["residual_add"]:
            x = x + x0
        flat = x.reshape(x.shape[0], -1) if self.c["flatten"] == "pos_major" else x.transpose(1, 2).reshape(x.shape[0], -1)
        return F.linear(flat, self.W["sig_dense/kernel"].t(), self.W["sig_dense/bias"])

    def _gate(self, z):
        if self.c["recurrent_activation"] == "hard_sigmoid":
            return torch.clamp(z * 0.2 + 0.5, 0.0, 1.0)
        return torch.sigmoid(z)

    def _lstm(self, x, layer, direction):
        W = self.W
        K, R, b = (W["%s/%s/%s" % (layer, direction, p)] for p in ("kernel", "recurrent_kernel", "bias"))
        n, T, _ = x.shape
        u = R.shape[0]
        h = x.new_zeros(n, u)
        c = x.new_zeros(n, u)
        ys = [None] * T
        steps = range(T) if direction == "forward" else reversed(range(T))
        order = self.c["gate_order"]
        for t in steps:
            z = torch.addmm(b, x[:, t], K) + h @ R
            parts = dict(zip(order, torch.split(z, u, dim=1)))
            i, f, o = self._gate(parts["i"]), self._gate(parts["f"]), self._gate(parts["o"])
            c = f * c + i * torch.tanh(parts["c"])
            h = o * torch.tanh(c)
            ys[t] = h
        return torch.stack(ys, dim=1)

    def _bidir(self, x, layer):
        return torch.cat([self._lstm(x, layer, "forward"), self._lstm(x, layer, "backward")], dim=-1)

    def predict(self, S, X):
        """S [n, W, 50], X [n, W, 6] (float64 tensors) -> softmax [n, n_class]: model.predict([S[..., None], X])"""
        n, Wn, _ = X.shape
        sf = self.signal_features(S.reshape(n * Wn, -1)).reshape(n, Wn, -1)
        if self.c["zero_signal"]:
            sf = torch.zeros_like(sf)
        r = self._bn(self._bidir(X, "read_rnn1"), "bn_read_rnn1")
        r = self._bn(self._bidir(r, "read_rnn11"), "bn_read_rnn11")
        tot = torch.cat([r, sf] if self.c["concat"] == "read_then_signal" else [sf, r], dim=-1)
        tot = self._bn(self._bidir(tot, "total_rnn1"), "bn_total_rnn1")
        tot = self._bidir(tot, "total_rnn2")
        d = tot
        for role in ("dense1", "dense2", "main_out"):
            d = torch.relu(F.linear(d, self.W[role + "/kernel"].t(), self.W[role + "/bias"]))
        flat = d.reshape(n, -1) if self.c["final_flatten"] == "t_major" else d.transpose(1, 2).reshape(n, -1)
        feat = torch.relu(F.linear(flat, self.W["feature/kernel"].t(), self.W["feature/bias"]))
        return torch.softmax(F.linear(feat, self.W["final_out/kernel"].t(), self.W["final_out/bias"]), dim=1)
