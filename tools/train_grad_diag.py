"""Diagnostic: per-tensor gradient error of the CUDA training path vs torch autograd in fp64 AND in fp32 (tells a kernel bug from
the conditioning of the clipped activations).  usage (GPU box): python tools/train_grad_diag.py [model index 0|1]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nanoreviser_b200 import train, weights  # noqa: E402
import test_train_gpu as T  # noqa: E402

which = int(sys.argv[1]) if len(sys.argv) > 1 else 1
w = weights.load_species("ecoli")[which]
rng = np.random.default_rng(7 + which)
B = 24
S, X, y, mask = T._inputs(rng, B, w.window, w.n_class)
tm = train.TrainModel(window=w.window, n_class=w.n_class, weights=w, seed=3)
tm.forward_backward(S, X, y, training=True, dropout_mask=mask)
dev = torch.device("cuda")
grads = {}
for dt in (torch.float64, torch.float32):
    P = {k: v.detach().to(dt).clone().requires_grad_(True) for k, v in tm.p.items()}
    zeros = torch.zeros
    T.torch_dtype = dt
    loss, *_ = T._ref_graph(torch, P, torch.tensor(S, dtype=dt, device=dev), torch.tensor(X, dtype=dt, device=dev), torch.tensor(y, device=dev),
                            torch.tensor(mask, dtype=dt, device=dev), None, w.n_class, keep=(keep := {}))
    loss.backward()
    if dt == torch.float64:
        Bq, Tq = B, w.window
        tmaj = lambda t: t.grad.permute(1, 0, 2).reshape(Tq * Bq, -1)            # [B,T,F] -> time-major rows
        for mine, ref in (("dout3", "out3"), ("dx3", "t1"), ("bnr2_dx", "out2"), ("dx2", "tot"), ("bnr1_dx", "out1")):
            a, r = tm._buf[mine].double(), tmaj(keep[ref])
            e = (a - r).abs()
            half = r.shape[1] // 2
            print("d(%s): max err / max %.3e   first half of the columns %.3e, second half %.3e" % (ref, float(e.max() / r.abs().max()),
                  float(e[:, :half].max() / r.abs().max()), float(e[:, half:].max() / r.abs().max())))
    grads[dt] = {k: p.grad.double().cpu().numpy() for k, p in P.items()}
g64 = grads[torch.float64]
g32 = grads[torch.float32]
print("%-10s %12s %12s %12s" % ("tensor", "cuda vs f64", "torch32 vs f64", "|g|max"))
for k in g64:
    a = tm.g[k].double().cpu().numpy()
    e = np.abs(a - g64[k]); i = np.unravel_index(e.argmax(), e.shape)
    e32 = np.abs(g32[k] - g64[k]).max() / np.abs(g64[k]).max()
    print("%-10s %12.3e %12.3e %12.3e  worst at %s: cuda %.6e ref %.6e ; rel L2 %.3e" % (k, e.max() / np.abs(g64[k]).max(), e32, np.abs(g64[k]).max(), i, a[i], g64[k][i],
          np.linalg.norm(a - g64[k]) / np.linalg.norm(g64[k])))
