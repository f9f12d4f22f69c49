"""Precision study on the CPU (not collected by pytest): what does the network's output lose when the two
2^-11-sized correction passes of the split-fp16 tensor-core products (x_lo*W_hi, x_hi*W_lo) run in 8-bit
float formats (tcgen05 kind::f8f6f4, twice the fp16 rate) instead of fp16?

Emulates operand rounding of every tensor-core product of the GPU path (read_rnn11, total_rnn1, total_rnn2,
dense1, dense2; BN of read_rnn11 / total_rnn1 folded into the next layer's kernel exactly as nrv_api.cu
does) with fp32 accumulation (torch CPU matmul) and compares the softmax output with the fp64 goldens of
the unitest reads.

    python tests/precision_study.py [--species ecoli human] [--reads 0 1 2 3 4] [--modes f16x3 e4m3 ...]

Test infrastructure: imports the oracle; nothing in the product imports this file.
"""
import argparse
import glob
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanoreviser_b200 import weights  # noqa: E402
from oracle import nanorev_oracle as orc  # noqa: E402

torch.set_num_threads(os.cpu_count() or 8)
F8 = {"e4m3": (torch.float8_e4m3fn, 448.0), "e5m2": (torch.float8_e5m2, 57344.0)}


def t32(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def f16(x):
    return x.to(torch.float16).to(torch.float32)


def q8(x, fmt, scale_log2):
    """round x * 2^s to the 8-bit float format (saturating), return the value it represents (scaled back)"""
    dt, mx = F8[fmt]
    s = 2.0 ** scale_log2
    return (x * s).clamp(-mx, mx).to(dt).to(torch.float32) / s


class Mode:
    """how one product x @ W is evaluated.  kinds: f32 | f16x3 | f16x1 | fp8 (fmt_xlo, fmt_whi, fmt_xhi, fmt_wlo)"""

    def __init__(self, name):
        self.name = name
        self.stats = {}

    def prep_w(self, W):
        W = t32(W)
        Wh = f16(W)
        Wl = f16(W - Wh)
        wmax = float(W.abs().max())
        # power-of-two scales that put the largest |W_hi| / |W_lo| just below the e4m3 maximum
        sw = int(np.floor(np.log2(448.0 / max(wmax, 1e-30))))
        swl = int(np.floor(np.log2(448.0 / max(float(Wl.abs().max()), 1e-30))))
        if W8_KERNEL_SCALES[0]:
            # the kernel's scales (nrv_api.cu pack_model, F8): one accumulator scale S = 7 + b for all passes with UNSCALED fp16
            # activations => W_hi8 = e4m3(W_hi 2^(b-5)) (meets x_lo 2^12), W_lo8 = e4m3(W_lo 2^(7+b)) (meets x_hi 2^0)
            swl = sw + 7
            sw = sw - 5
        return dict(W=W, Wh=Wh, Wl=Wl, sw=sw, swl=swl,
                    Wh8=q8(Wh, "e4m3", sw), Wl8=q8(Wl, "e4m3", swl),
                    Wh8_52=q8(Wh, "e5m2", sw), Wl8_52=q8(Wl, "e5m2", swl))

    def mm(self, x, w, xs_hi=8, xs_lo=19):
        """x fp32 activations [.., K]; xs_hi / xs_lo: power-of-two scales of the 8-bit copies of x_hi / x_lo"""
        n = self.name
        if n == "f32":
            return x @ w["W"]
        xh = f16(x)
        xl = f16(x - xh)
        if n == "f16x3":
            return xh @ w["Wh"] + (xl @ w["Wh"] + xh @ w["Wl"])
        if n == "f16x1":
            return xh @ w["Wh"]
        if n == "f16x2w":      # weights exact to 22 bits, activations single fp16
            return xh @ w["Wh"] + xh @ w["Wl"]
        if n == "e4m3":        # both correction passes in e4m3 x e4m3
            return xh @ w["Wh"] + (q8(xl, "e4m3", xs_lo) @ w["Wh8"] + q8(xh, "e4m3", xs_hi) @ w["Wl8"])
        if n == "e4m3_lo3":    # x_lo in e5m2 (3 significant bits), everything else e4m3
            return xh @ w["Wh"] + (q8(xl, "e5m2", xs_lo) @ w["Wh8"] + q8(xh, "e4m3", xs_hi) @ w["Wl8"])
        if n == "e5m2w":       # weights of the correction passes in e5m2
            return xh @ w["Wh"] + (q8(xl, "e4m3", xs_lo) @ w["Wh8_52"] + q8(xh, "e4m3", xs_hi) @ w["Wl8_52"])
        if n == "half8":       # x_hi*W_lo stays fp16 (2.5 passes)
            return xh @ w["Wh"] + (q8(xl, "e4m3", xs_lo) @ w["Wh8"] + xh @ w["Wl"])
        raise ValueError(n)


def hard_sigmoid(x):
    return (0.2 * x + 0.5).clamp(0.0, 1.0)


def bn_affine(bn):
    g, b, m, v = (bn[i].astype(np.float64) for i in range(4))
    inv = g / np.sqrt(v + orc.BN_EPS)
    return inv, b - m * inv


BASE_MODE = ["f32"]
W8_KERNEL_SCALES = [False]
HEADS_D1_ONLY = [False]
H_SCALES = [8, 19]     # log2 scales of the 8-bit copies of h (hi, lo)
PROJ_F16_LAYERS = []   # --proj-f16-layers: layers whose PROJECTION stays f16x3 while their recurrence runs the emulated mode
PROJ_SPLIT = {}        # --proj-split LAYER COLS: that layer's projection runs the emulated mode on its first COLS input columns only (the
                       # raw h of the layer below) and f16x3 on the rest (total_rnn1: the 64 CNN-feature columns, which are unbounded)
REC_MODE = [None]      # --rec-mode: how the recurrent product h @ Wr of the emulated layers is evaluated (default: same as the projection)


def _rows(w, a, b):
    return {k: (v[a:b] if hasattr(v, "shape") and getattr(v, "ndim", 0) == 2 else v) for k, v in w.items()}


def lstm_dir(mode, x, wk, wr, bias, reverse, xs_hi, xs_lo, proj_f16=False, split=0):
    """x [B, T, in] (fp32 torch) -> [B, T, u]; h is in [-1, 1]: 8-bit copies scaled by 2^8 (hi) / 2^19 (lo)"""
    B, T, _ = x.shape
    u = wr["W"].shape[0]
    pm = Mode("f16x3") if (proj_f16 and mode.name != "f32") else mode
    x2 = x.reshape(B * T, -1)
    if split and mode.name not in ("f32", "f16x3"):
        zin = (mode.mm(x2[:, :split], _rows(wk, 0, split), H_SCALES[0], H_SCALES[1]) +
               Mode("f16x3").mm(x2[:, split:], _rows(wk, split, x2.shape[1]))).reshape(B, T, 4 * u) + bias
    else:
        zin = pm.mm(x2, wk, xs_hi, xs_lo).reshape(B, T, 4 * u) + bias
    if REC_MODE[0] and mode.name != "f32":
        mode = Mode(REC_MODE[0])
    h = torch.zeros(B, u)
    c = torch.zeros(B, u)
    out = torch.empty(B, T, u)
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        z = zin[:, t] + (mode.mm(h, wr, H_SCALES[0], H_SCALES[1]) if (t != (T - 1 if reverse else 0)) else 0.0)
        i, f = hard_sigmoid(z[:, :u]), hard_sigmoid(z[:, u:2 * u])
        g, o = torch.tanh(z[:, 2 * u:3 * u]), hard_sigmoid(z[:, 3 * u:])
        c = f * c + i * g
        h = o * torch.tanh(c)
        out[:, t] = h
    return out


class Net:
    """one model with the GPU path's operand preparation (BN folds) under a given Mode"""

    def __init__(self, m, mode, tensor_layers=(1, 2, 3), tensor_heads=True):
        self.m, self.mode = m, mode
        f32m = Mode("f32")
        self.layers = []
        for li in range(4):
            dirs = []
            for d in m.lstm[li]:
                wk = d.kernel.astype(np.float64)
                b = d.bias.astype(np.float64)
                if li in (2, 3):                                   # BN of the previous layer folded into Wk / bias
                    inv, off = bn_affine(m.bn_rnn[li - 1])
                    if li == 2:                                    # [read_rnn11 (BN) | CNN features (no BN)]
                        inv = np.concatenate([inv, np.ones(64)])
                        off = np.concatenate([off, np.zeros(64)])
                    b = b + off @ wk
                    wk = wk * inv[:, None]
                md = mode if li in tensor_layers else (Mode(BASE_MODE[0]) if li >= 1 else f32m)
                dirs.append((md, md.prep_w(wk), md.prep_w(d.recurrent), t32(b)))
            self.layers.append(dirs)
        hm = mode if tensor_heads else Mode(BASE_MODE[0])
        self.hm = hm
        self.d1 = hm.prep_w(m.dense1_k)
        self.hm2 = Mode("f16x3") if (HEADS_D1_ONLY[0] and hm.name != "f32") else hm      # --heads-d1-only: dense2 stays f16x3
        self.d2 = self.hm2.prep_w(m.dense2_k)

    def bilstm(self, li, x, xs_hi=8, xs_lo=19):
        outs = []
        for k, (md, wk, wr, b) in enumerate(self.layers[li]):
            outs.append(lstm_dir(md, x, wk, wr, b, k == 1, xs_hi, xs_lo, proj_f16=li in PROJ_F16_LAYERS, split=PROJ_SPLIT.get(li, 0)))
        return torch.cat(outs, dim=-1)

    def forward(self, sig_feat, X):
        m = self.m
        inv, off = bn_affine(m.bn_rnn[0])
        r1 = self.bilstm(0, X) * t32(inv) + t32(off)               # read_rnn1 fp32 SIMT; its BN is applied, not folded
        amax = float(r1.abs().max())
        # 8-bit copies of a1: static power-of-two scale from the observed range (the kernel would use the BN's own bound)
        s_hi = int(np.floor(np.log2(448.0 / max(amax, 1e-9))))
        r2 = self.bilstm(1, r1, s_hi, s_hi + 11)
        tot = torch.cat([r2, sig_feat], dim=-1)
        fmax = float(tot.abs().max())
        s_hi = int(np.floor(np.log2(448.0 / max(fmax, 1.0))))
        t1 = self.bilstm(2, tot, *((H_SCALES[0], H_SCALES[1]) if W8_KERNEL_SCALES[0] else (s_hi, s_hi + 11)))
        t2 = self.bilstm(3, t1, H_SCALES[0], H_SCALES[1])
        B, T, _ = t2.shape
        d = torch.relu(self.hm.mm(t2.reshape(B * T, -1), self.d1, H_SCALES[0], H_SCALES[1]) + t32(m.dense1_b))
        dmax = float(d.abs().max())
        s_hi = int(np.floor(np.log2(448.0 / max(dmax, 1.0))))
        d = torch.relu(self.hm2.mm(d, self.d2, s_hi, s_hi + 11) + t32(m.dense2_b))
        d = torch.relu(d @ t32(m.main_k) + t32(m.main_b)).reshape(B, -1)
        feat = torch.relu(d @ t32(m.feat_k) + t32(m.feat_b))
        logits = feat @ t32(m.final_k) + t32(m.final_b)
        return torch.softmax(logits.double(), dim=1).numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--species", nargs="+", default=["ecoli", "human"])
    ap.add_argument("--reads", nargs="+", type=int, default=[0, 1, 2, 3, 4])
    ap.add_argument("--modes", nargs="+", default=["f16x3", "e4m3", "e4m3_lo3", "e5m2w", "half8", "f16x2w", "f16x1"])
    ap.add_argument("--layers", nargs="+", type=int, default=[1, 2, 3], help="LSTM layers on the emulated tensor path")
    ap.add_argument("--max-windows", type=int, default=0)
    ap.add_argument("--no-heads", action="store_true", help="dense heads in fp32")
    ap.add_argument("--rec-mode", default=None, help="mode of the recurrent products (e.g. f16x3 while the projections use e4m3)")
    ap.add_argument("--proj-f16-layers", nargs="*", type=int, default=[], help="layers whose projection stays f16x3 (only their recurrence is emulated)")
    ap.add_argument("--h-scales", nargs=2, type=int, default=[8, 19], help="log2 scales of the 8-bit copies of h: hi, lo (default 8 19; the unified format is 1 12)")
    ap.add_argument("--heads-d1-only", action="store_true", help="of the dense head only the first dense (input h of total_rnn2) runs the emulated mode")
    ap.add_argument("--kernel-scales", action="store_true", help="8-bit weight scales as pack_model derives them (S = 7 + b, activations unscaled)")
    ap.add_argument("--proj-split", nargs=2, type=int, default=None, metavar=("LAYER", "COLS"),
                    help="that layer's projection: emulated mode on the first COLS input columns, f16x3 on the rest (overrides --proj-f16-layers for it)")
    ap.add_argument("--base-mode", default="f16x3", help="mode of the tensor layers NOT listed in --layers (the GPU default is f16x3)")
    a = ap.parse_args()
    REC_MODE[0] = a.rec_mode
    BASE_MODE[0] = a.base_mode
    PROJ_F16_LAYERS[:] = a.proj_f16_layers
    if a.proj_split:
        PROJ_SPLIT[a.proj_split[0]] = a.proj_split[1]
    W8_KERNEL_SCALES[0] = a.kernel_scales
    HEADS_D1_ONLY[0] = a.heads_d1_only
    H_SCALES[:] = a.h_scales
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "fast5", "*.fast5")))
    for sp in a.species:
        m1, m2 = weights.load_species(sp, os.path.join(ROOT, "model"))
        gold = np.load(os.path.join(ROOT, "tests", "golden", "forward_%s.npz" % sp))
        res = {md: dict(d=[0.0, 0.0], flips=[0, 0], n=0, seq=0) for md in a.modes}
        for k in a.reads:
            a0, starts, length, bases, signal, evm, evs = orc.get_read_data(files[k])
            sig = np.asarray(signal)[a0:]
            win, mean, std, shift, scale = orc.signal_segmentation(sig, starts, length[-1])
            x = orc.feature_columns(bases, mean, std, shift, scale, length, evm, evs)
            N = len(x)
            M = N - m1.window
            if a.max_windows:
                M = min(M, a.max_windows)
            idx = np.arange(M)[:, None] + np.arange(m1.window)[None, :]
            for md in a.modes:
                ys = []
                for mi, m in enumerate((m1, m2)):
                    sf = t32(orc.cnn_branch(m, win, np.float32))
                    net = Net(m, Mode(md), tensor_layers=tuple(a.layers), tensor_heads=not a.no_heads)
                    P = np.concatenate([net.forward(sf[idx[s:s + 8192]], t32(x)[idx[s:s + 8192]]) for s in range(0, M, 8192)])
                    G = gold["r%d_P%d_f64" % (k, mi + 1)][:M]
                    r = res[md]
                    r["d"][mi] = max(r["d"][mi], float(np.abs(P - G).max()))
                    y = P.argmax(1)
                    r["flips"][mi] += int((y != gold["r%d_y%d_f64" % (k, mi + 1)][:M]).sum())
                    ys.append(y)
                res[md]["n"] += M
                if not a.max_windows:
                    core = orc.get_base_1(bases[5:5 + M], ys[0], ys[1] + 2)
                    rev = "".join(bases[:5]) + core + "".join(bases[5 + M:])
                    res[md]["seq"] += int(rev == gold["r%d_revised" % k].tobytes().decode())
            print("[%s] read %d done (%d windows)" % (sp, k, M), flush=True)
        for md in a.modes:
            r = res[md]
            print("%-6s %-9s max|dP1| %.2e  max|dP2| %.2e  flips %d + %d of %d  identical sequences %d/%d" % (
                sp, md, r["d"][0], r["d"][1], r["flips"][0], r["flips"][1], r["n"], r["seq"], len(a.reads)), flush=True)


if __name__ == "__main__":
    main()
