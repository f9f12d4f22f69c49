"""Diagnostic: layer 2 backward on the REAL activations of a training step vs autograd on the same tensors."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nanoreviser_b200 import train, weights
import test_train_gpu as TT
w = weights.load_species("ecoli")[1]
rng = np.random.default_rng(8)
B, T = 24, w.window
S, X, y, mask = TT._inputs(rng, B, T, w.n_class)
tm = train.TrainModel(window=T, n_class=w.n_class, weights=w, seed=3)
tm.forward_backward(S, X, y, training=True, dropout_mask=mask)
l, u, n_in = 2, 128, 192
tot = tm._buf["tot"].clone(); dout = tm._buf["bnr2_dx"].clone()
dx_full = tm._buf["dx2"].clone()
g_full = {k: tm.g[k].clone() for k in tm.g if k.startswith("l2")}
out = tm._lstm_fwd(l, tot, n_in, n_in, T, B)
dx = tm._lstm_bwd(l, tot, n_in, n_in, dout, T, B)
torch.cuda.synchronize()
print("re-run equals the in-graph run: dx %s" % torch.equal(dx, dx_full), {k: torch.equal(tm.g[k], v) for k, v in g_full.items()})
dev = tm.dev
P = {k: tm.p[k].detach().double().clone().requires_grad_(True) for k in tm.p if k.startswith("l2")}
Xr = tot.double().reshape(T, B, n_in).clone().requires_grad_(True)
hs = lambda z: torch.clamp(0.2 * z + 0.5, 0.0, 1.0)
outs = []
zs = {}
for d in range(2):
    Wk, Wr, b = P["l2%d_k" % d], P["l2%d_r" % d], P["l2%d_b" % d]
    h = torch.zeros(B, u, dtype=torch.float64, device=dev); c = torch.zeros_like(h)
    o = [None] * T
    for t in (range(T) if d == 0 else range(T - 1, -1, -1)):
        z = Xr[t] @ Wk + b + h @ Wr
        zs[(d, t)] = z.detach()
        i, f, g, og = hs(z[:, :u]), hs(z[:, u:2 * u]), torch.tanh(z[:, 2 * u:3 * u]), hs(z[:, 3 * u:])
        c = f * c + i * g
        h = og * torch.tanh(c)
        o[t] = h
    outs.append(torch.stack(o, 0))
O = torch.cat(outs, -1)
(O * dout.double().reshape(T, B, 2 * u)).sum().backward()
print("fwd", float((out.double().reshape(T, B, 2 * u) - O).abs().max()))
print("dx ", float((dx.double().reshape(T, B, n_in) - Xr.grad).abs().max() / Xr.grad.abs().max()))
for k, p in P.items():
    print(k, float((tm.g[k].double() - p.grad).abs().max() / p.grad.abs().max()))
# how close do pre-activations come to the kinks of hard_sigmoid (|z| = 2.5)?
for d in range(2):
    zz = torch.stack([zs[(d, t)] for t in range(T)])
    gates = torch.cat([zz[..., :u], zz[..., u:2 * u], zz[..., 3 * u:]], -1)
    dist = (gates.abs() - 2.5).abs()
    print("dir %d: min distance of an i/f/o pre-activation to +-2.5: %.3e ; count within 1e-5: %d, 1e-4: %d" % (d, float(dist.min()), int((dist < 1e-5).sum()), int((dist < 1e-4).sum())))
