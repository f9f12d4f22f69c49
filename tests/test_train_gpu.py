"""Training path (nanoreviser_b200/train.py + csrc/nrv_train.cu) on the GPU.

Oracle for this floating-point path: an fp64 torch autograd graph of the train model of nanorevutils/lstmmodel.py:32-81 written
here from torch primitives (conv1d, manual LSTM loop with hard_sigmoid, batch statistics, the given dropout mask, softmax
cross-entropy + 0.4 x centre loss).  Tolerance: every parameter gradient within 1e-3 of the largest reference gradient entry of
that tensor (fp32 SIMT kernels against fp64; measured ~1e-5).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-3


def _ref_graph(torch, P, S, X, y, mask, cw, n_class, keep=None, gate_masks=None):
    """P: dict of fp64 leaf tensors named like TrainModel.p -> (loss, ce, l2, probs, batch statistics)"""
    F = torch.nn.functional
    B, T = X.shape[0], X.shape[1]
    hs = lambda z: torch.clamp(0.2 * z + 0.5, 0.0, 1.0)
    stats = {}

    def bn(name, x, dims, shape):
        mean = x.mean(dim=dims, keepdim=True)
        var = ((x - mean) ** 2).mean(dim=dims, keepdim=True)
        stats[name] = (mean.reshape(-1).detach(), var.reshape(-1).detach())
        return P[name + "_g"].reshape(shape) * (x - mean) / torch.sqrt(var + 1e-3) + P[name + "_b"].reshape(shape)

    s = S.reshape(B * T, 1, 50)
    c1 = torch.relu(F.conv1d(s, P["conv1_k"].permute(2, 1, 0), P["conv1_b"], padding=1))
    b1 = bn("bn1", c1, (0, 2), (1, 8, 1))
    c2 = torch.relu(F.conv1d(b1, P["conv2_k"].permute(2, 1, 0), P["conv2_b"], padding=1))
    res = bn("bn2", c2, (0, 2), (1, 8, 1)) + s
    res = res * mask.reshape(B * T, 50, 8).permute(0, 2, 1) / 0.8
    flat = res.permute(0, 2, 1).reshape(B * T, 400)                      # index pos*8 + ch
    sig = (flat @ P["sig_k"] + P["sig_b"]).reshape(B, T, 64)

    def lstm_dir(x, l, d):
        Wk, Wr, b = P["l%d%d_k" % (l, d)], P["l%d%d_r" % (l, d)], P["l%d%d_b" % (l, d)]
        u = Wr.shape[0]
        h = torch.zeros(B, u, dtype=x.dtype, device=x.device); c = torch.zeros_like(h)
        out = [None] * T
        for t in (range(T) if d == 0 else range(T - 1, -1, -1)):
            z = x[:, t] @ Wk + b + h @ Wr
            gates = torch.cat([z[:, :2 * u], z[:, 3 * u:]], dim=1).detach()
            stats["_kink"] = min(stats.get("_kink", 1e9), float(((gates.abs() - 2.5).abs()).min()))
            if gate_masks is None:
                i, f, o = hs(z[:, :u]), hs(z[:, u:2 * u]), hs(z[:, 3 * u:])
            else:       # same value, but the derivative (0.2 inside the linear range, 0 outside) taken where the CUDA path took it
                gm = gate_masks[(l, d)][t].to(z.dtype)
                hm = lambda zz, mm: hs(zz).detach() + 0.2 * mm * (zz - zz.detach())
                i, f, o = hm(z[:, :u], gm[:, :u]), hm(z[:, u:2 * u], gm[:, u:2 * u]), hm(z[:, 3 * u:], gm[:, 3 * u:])
            g = torch.tanh(z[:, 2 * u:3 * u])
            c = f * c + i * g
            h = o * torch.tanh(c)
            out[t] = h
        return torch.stack(out, dim=1)

    def kept(name, t):                  # diagnostics (tools/train_grad_diag.py): intermediate tensors with their gradients
        if keep is not None:
            t.retain_grad()
            keep[name] = t
        return t

    bil = lambda x, l: torch.cat([lstm_dir(x, l, 0), lstm_dir(x, l, 1)], dim=-1)
    r1 = bn("bnr0", kept("out0", bil(X, 0)), (0, 1), (1, 1, 32))
    r2 = bn("bnr1", kept("out1", bil(r1, 1)), (0, 1), (1, 1, 128))
    tot = kept("tot", torch.cat([r2, sig], dim=-1))
    t1 = kept("t1", bn("bnr2", kept("out2", bil(tot, 2)), (0, 1), (1, 1, 256)))
    t2 = kept("out3", bil(t1, 3))
    d = torch.relu(t2 @ P["d1_k"] + P["d1_b"])
    d = torch.relu(d @ P["d2_k"] + P["d2_b"])
    d = torch.relu(d @ P["m_k"] + P["m_b"])
    feat = torch.relu(d.reshape(B, T * 6) @ P["f_k"] + P["f_b"])
    logits = feat @ P["o_k"] + P["o_b"]
    logp = torch.log_softmax(logits, dim=1)
    w = cw[y] if cw is not None else torch.ones(B, dtype=X.dtype, device=X.device)
    ce = -(w * logp[torch.arange(B), y]).mean()
    l2_i = ((feat - P["centers"][y]) ** 2).sum(dim=1)
    l2 = l2_i.mean()
    stats["_l2_acc"] = float((l2_i.detach() <= 0.5).double().mean())
    return ce + 0.4 * l2, ce, l2, torch.softmax(logits, dim=1), stats


def _inputs(rng, B, T, n_class):
    S = rng.normal(0.0, 1.0, (B, T, 50)).astype(np.float32)
    X = np.stack([rng.choice([250, 180, 100, 30], (B, T)) / 300.0, rng.normal(1.0, 0.1, (B, T)), rng.gamma(2.0, 0.1, (B, T)),
                  rng.choice([0.2, 0.3, 0.5, 1.0, 1.5], (B, T)), rng.normal(100, 15, (B, T)), rng.gamma(2.0, 1.0, (B, T))], axis=-1).astype(np.float32)
    y = rng.integers(0, n_class, B)
    mask = rng.random((B, T, 50, 8)) >= 0.2
    return S, X, y, mask


@pytest.mark.parametrize("which,cw_mode", [(0, "applied"), (1, "keras")])
def test_gradients_match_fp64_autograd(weights_by_species, which, cw_mode):
    """Whole train graph, one batch, shipped ecoli weights as the starting point: loss, probabilities, batch statistics and the
    gradient of EVERY parameter against fp64 autograd.
    hard_sigmoid has kinks at |z| = 2.5 and most gates of the trained networks are saturated, so the few unsaturated elements carry
    the gradient: ONE pre-activation within fp32 rounding of a kink (seen: 3e-7 away, tools/train_lstm_real_diag.py) falls on the
    other side of it in fp64 and moves a whole layer's gradient by percents although both sides are right.  The fp64 graph
    therefore takes each gate's derivative (0.2 or 0) where the CUDA path took it -- read back from its stored gate activations --
    while every VALUE stays its own; how close the batch came to a kink is printed."""
    import torch
    from nanoreviser_b200 import train
    w = weights_by_species("ecoli")[which]
    B, T = 24, w.window
    class_weight = {0: 3, 1: 5, 2: 1, 3: 1, 4: 1, 5: 1}
    tm = train.TrainModel(window=T, n_class=w.n_class, weights=w, seed=3, class_weight_mode=cw_mode)
    dev = torch.device("cuda")
    cw = torch.tensor([class_weight[k] for k in range(w.n_class)], dtype=torch.float64, device=dev) if cw_mode == "applied" else None
    S, X, y, mask = _inputs(np.random.default_rng(7 + which), B, T, w.n_class)
    m = tm.forward_backward(S, X, y, class_weight=class_weight, training=True, dropout_mask=mask)
    torch.cuda.synchronize()
    gate_masks = {}
    for l in range(4):
        for d in range(2):
            a = tm._buf["z%d%d" % (l, d)].reshape(T, B, -1)                 # activated gates, time-major
            gate_masks[(l, d)] = (a > 0) & (a < 1)
    P = {k: v.detach().double().clone().requires_grad_(True) for k, v in tm.p.items()}
    loss, ce, l2, probs, stats = _ref_graph(torch, P, torch.tensor(S, dtype=torch.float64, device=dev), torch.tensor(X, dtype=torch.float64, device=dev),
                                            torch.tensor(y, device=dev), torch.tensor(mask, dtype=torch.float64, device=dev), cw, w.n_class,
                                            gate_masks=gate_masks)
    print("closest gate pre-activation to a hard_sigmoid kink: %.2e" % stats.pop("_kink"))
    assert abs(m["l2_loss1_acc"] - stats.pop("_l2_acc")) <= 1.0 / B + 1e-9
    loss.backward()
    loss, ce, l2 = float(loss.detach()), float(ce.detach()), float(l2.detach())
    assert abs(m["loss"] - loss) <= 1e-4 * max(1.0, abs(loss)), (m, loss)
    assert abs(m["final_out_loss"] - ce) <= 1e-4 * max(1.0, abs(ce)) and abs(m["l2_loss1_loss"] - l2) <= 1e-4 * max(1.0, l2)
    assert np.abs(tm._buf["probs"].cpu().numpy() - probs.detach().cpu().numpy()).max() <= 1e-5
    for name, (mean, var) in stats.items():                      # batch statistics feed the moving averages
        assert np.allclose(tm._buf[name + "_bm"].cpu().numpy(), mean.cpu().numpy(), rtol=1e-4, atol=1e-5), name
        assert np.allclose(tm._buf[name + "_bv"].cpu().numpy(), var.cpu().numpy(), rtol=1e-4, atol=1e-6), name
    g_all = max(float(p.grad.abs().max()) for p in P.values())
    worst = {}
    for k, p in P.items():
        g_ref = p.grad.cpu().numpy()
        g = tm.g[k].cpu().numpy().astype(np.float64)
        assert np.abs(g_ref).max() > 0, k                          # every parameter takes part
        # relative to the tensor's largest gradient; tensors whose whole gradient is below fp32 resolution of the step (a fully
        # saturated layer: 1e-10) are held to the global scale instead
        worst[k] = float(np.abs(g - g_ref).max() / max(np.abs(g_ref).max(), 1e-6 * g_all))
    bad = {k: v for k, v in worst.items() if v > GRAD_TOL}
    assert not bad, (sorted(bad.items(), key=lambda kv: -kv[1]), max(worst.values()))


def test_adam_and_moving_average_updates(weights_by_species):
    import torch
    from nanoreviser_b200 import train
    w = weights_by_species("ecoli")[0]
    rng = np.random.default_rng(17)
    S, X, y, mask = _inputs(rng, 16, w.window, w.n_class)
    tm = train.TrainModel(window=w.window, n_class=w.n_class, weights=w, seed=1)
    p0 = {k: v.detach().double().cpu().numpy().copy() for k, v in tm.p.items()}
    s0 = {k: v.detach().double().cpu().numpy().copy() for k, v in tm.s.items()}
    mom, vel = {k: np.zeros_like(v) for k, v in p0.items()}, {k: np.zeros_like(v) for k, v in p0.items()}
    for step in (1, 2, 3):
        tm.train_on_batch(S, X, y, dropout_mask=mask)
        torch.cuda.synchronize()
        lr_t = 1e-3 * np.sqrt(1 - 0.999 ** step) / (1 - 0.9 ** step)
        for k in p0:
            g = tm.g[k].double().cpu().numpy()
            mom[k] = 0.9 * mom[k] + 0.1 * g
            vel[k] = 0.999 * vel[k] + 0.001 * g * g
            p0[k] = p0[k] - lr_t * mom[k] / (np.sqrt(vel[k]) + 1e-7)
            assert np.allclose(tm.p[k].cpu().numpy(), p0[k], rtol=2e-5, atol=2e-7), (step, k)
        for name, rows in tm._last_rows.items():
            s0[name + "_mean"] = 0.99 * s0[name + "_mean"] + 0.01 * tm._buf[name + "_bm"].double().cpu().numpy()
            s0[name + "_var"] = 0.99 * s0[name + "_var"] + 0.01 * tm._buf[name + "_bv"].double().cpu().numpy() * rows / (rows - 1.001)
            assert np.allclose(tm.s[name + "_mean"].cpu().numpy(), s0[name + "_mean"], rtol=1e-5, atol=1e-6), (step, name)
            assert np.allclose(tm.s[name + "_var"].cpu().numpy(), s0[name + "_var"], rtol=1e-5, atol=1e-6), (step, name)


def test_graph_replay_equals_eager_steps(weights_by_species):
    """train_on_batch(graph=True) -- the step captured into a CUDA graph on its third call and replayed afterwards, with a
    validation pass of another batch size in between -- leaves the same parameters as the eager launches (same seeds, same masks:
    the dropout mask is a function of (seed, step, element) read from device scalars).  "The same" is to fp32 rounding, not bitwise:
    the split-K products accumulate with atomics, whose order differs from run to run (eager against eager as well)."""
    import torch
    from nanoreviser_b200 import train
    w = weights_by_species("ecoli")[0]
    rng = np.random.default_rng(23)
    batches = [_inputs(rng, 32, w.window, w.n_class)[:3] for _ in range(6)]
    Sv, Xv, yv, _ = _inputs(rng, 20, w.window, w.n_class)
    models = []
    for graph in (False, True):
        tm = train.TrainModel(window=w.window, n_class=w.n_class, weights=w, seed=4)
        losses = []
        for k, (S, X, y) in enumerate(batches):
            losses.append(tm.train_on_batch(S, X, y, graph=graph)["loss"])
            if k == 3:
                tm.forward_backward(Sv, Xv, yv, training=False)           # other shapes through the same named buffers
        torch.cuda.synchronize()
        models.append((tm, losses))
    (a, la), (b, lb) = models
    assert "graph" in next(iter(b._graphs.values())) and not a._graphs
    assert np.allclose(la, lb, rtol=1e-4), (la, lb)
    for k in a.p:
        assert np.allclose(a.p[k].cpu().numpy(), b.p[k].cpu().numpy(), rtol=1e-4, atol=2e-6), k
    for k in a.s:
        assert np.allclose(a.s[k].cpu().numpy(), b.s[k].cpu().numpy(), rtol=1e-4, atol=2e-6), k


def test_deterministic_mode_is_bitwise_reproducible():
    """NRV_TRAIN_DETERMINISTIC=1 (no fp32 atomics: no split-K, whole-column sums, the centre gradient by one writer per entry):
    two processes with the same seeds end with bit-identical parameters and moving statistics after five steps."""
    import hashlib
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, hashlib, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from nanoreviser_b200 import train\n"
        "import test_train_gpu as T\n"
        "rng = np.random.default_rng(2)\n"
        "tm = train.TrainModel(window=11, n_class=6, seed=9)\n"
        "for k in range(5):\n"
        "    S, X, y, _ = T._inputs(rng, 48, 11, 6); X[..., 4:6] /= 100.0\n"
        "    tm.train_on_batch(S, X, y, class_weight={0: 3, 1: 5}, graph=(k >= 1))\n"
        "h = hashlib.sha256()\n"
        "for d in (tm.p, tm.s):\n"
        "    for k in sorted(d): h.update(d[k].cpu().numpy().tobytes())\n"
        "print('HASH', h.hexdigest())\n") % (root, os.path.join(root, "tests"))
    env = dict(os.environ, NRV_TRAIN_DETERMINISTIC="1")
    outs = []
    for _ in range(2):
        p = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append([l for l in p.stdout.splitlines() if l.startswith("HASH")][0])
    assert outs[0] == outs[1], outs


def test_fit_learns_and_the_saved_weights_drive_the_inference_engine(tmp_path):
    """Model.fit semantics end to end on a learnable synthetic task (the label is a function of the centre base's colour column),
    then the Keras-layout weight file goes through weights.load_model_weights into the inference engine, whose probabilities
    must agree with the training path's own validation graph."""
    from nanoreviser_b200 import engine, train, weights
    rng = np.random.default_rng(5)
    EPOCHS = 25
    T, N = 11, 1536
    # a learnable task shaped like the real one: every base has a signal level, the label is the centre base
    _, X, _, _ = _inputs(rng, N, T, 6)
    base = rng.integers(0, 4, (N, T))
    X[..., 0] = np.array([250, 180, 100, 30])[base] / 300.0
    X[..., 4:6] /= 100.0        # (raw event means ~100 saturate every gate of a freshly initialised read_rnn1)
    S = (np.array([-1.0, -0.3, 0.3, 1.0])[base][..., None] + 0.1 * rng.normal(size=(N, T, 50))).astype(np.float32)
    y = np.array([5, 4, 3, 2])[base[:, T // 2]]
    hist = {}
    models = []
    for n_class, labels in ((6, y), (5, y - 1)):
        tm = train.TrainModel(window=T, n_class=n_class, seed=11)
        h = tm.fit([S[..., None], X, labels.reshape(-1, 1)], [labels, np.zeros((N, 1))], class_weight={0: 3, 1: 5, 2: 1, 3: 1, 4: 1, 5: 1},
                   validation_split=0.125, shuffle=True, epochs=EPOCHS, batch_size=64, verbose=0)
        assert set(h) == {"loss", "final_out_loss", "l2_loss1_loss", "final_out_acc", "l2_loss1_acc", "val_loss", "val_final_out_loss",
                          "val_l2_loss1_loss", "val_final_out_acc", "val_l2_loss1_acc"} and all(len(v) == EPOCHS for v in h.values())
        assert 0.0 <= h["l2_loss1_acc"][-1] <= 1.0
        print("model with %d classes: loss %.3f -> %.3f, acc %.3f -> %.3f, val_acc %.3f -> %.3f" % (
            n_class, h["loss"][0], h["loss"][-1], h["final_out_acc"][0], h["final_out_acc"][-1], h["val_final_out_acc"][0], h["val_final_out_acc"][-1]))
        # 525 Adam steps at Keras' default rate from a random initialisation: the task is learned, on held-out windows too
        assert h["loss"][-1] < 0.8 * h["loss"][0] and h["final_out_acc"][-1] > h["final_out_acc"][0] + 0.2, h
        assert h["val_final_out_acc"][-1] > h["val_final_out_acc"][0] + 0.15, h
        fn = str(tmp_path / ("model%d.h5" % n_class))
        tm.save_weights(fn)
        w = weights.load_model_weights(fn)
        assert w.window == T and w.n_class == n_class
        got = tm.get_weights()
        assert np.array_equal(w.lstm[2][1].recurrent, got.lstm[2][1].recurrent) and np.array_equal(w.bn_rnn[1], got.bn_rnn[1])
        models.append((tm, w))
        hist[n_class] = h
    (tm1, w1), (tm2, w2) = models
    with engine.Reviser(w1, w2) as rv:
        p1, p2 = rv.predict_windows(S[:300], X[:300])
    assert np.abs(p1 - tm1.predict(S[:300], X[:300])).max() <= 2e-3
    assert np.abs(p2 - tm2.predict(S[:300], X[:300])).max() <= 2e-3


def test_training_fails_loudly_on_bad_input(weights_by_species):
    from nanoreviser_b200 import train
    tm = train.TrainModel(window=11, n_class=6)
    with pytest.raises(ValueError):
        tm.forward_backward(np.zeros((4, 11, 50), np.float32), np.zeros((4, 11, 6), np.float32), np.array([0, 1, 2, 6]))
    with pytest.raises(ValueError):
        tm.forward_backward(np.zeros((4, 13, 50), np.float32), np.zeros((4, 13, 6), np.float32), np.zeros(4))


def _synthetic_alignment(rng, bases: str):
    """A genome made FROM the read by random edits, so the true alignment (CIGAR) is known: -> (genome sequence, sam record)"""
    ref, cigar = [], []
    def push(op):
        if cigar and cigar[-1][1] == op:
            cigar[-1][0] += 1
        else:
            cigar.append([1, op])
    n, soft = len(bases), 3
    for i in range(soft, n):
        r = rng.random()
        edge = i < soft + 5 or i >= n - 5                  # match at both ends
        if edge or r < 0.90:
            ref.append(bases[i] if edge or rng.random() < 0.95 else "ACGT"[(("ACGT".index(bases[i])) + 1) % 4]); push("M")
        elif r < 0.95:
            push("I")                                       # base only in the read
        else:
            ref.append(bases[i]); push("M"); ref.append("ACGT"[int(rng.integers(0, 4))]); push("D")
    prefix = "".join(rng.choice(list("ACGT"), 40))
    genome = prefix + "".join(ref) + "".join(rng.choice(list("ACGT"), 40))
    sam = "\t".join(["read", "0", "chr1", str(len(prefix) + 1), "40", "%dS" % soft + "".join("%d%s" % (k, op) for k, op in cigar), "*", "0", "0", bases, "*"])
    return genome, sam


def test_cli_trains_both_models_from_fast5_and_sam(tmp_path, fast5_files):
    """NanoReviser_train.py end to end: fast5 + SAM (synthetic alignments with known CIGARs; the mapper itself is an external program)
    -> per-read .npz -> window tensors -> two fits -> Keras-layout weight files that the inference loader accepts, history / parameter
    files with the reference's names."""
    import subprocess
    import sys
    from nanoreviser_b200 import fast5, weights
    rng = np.random.default_rng(3)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    f5dir, samdir = tmp_path / "fast5", tmp_path / "sam"
    f5dir.mkdir(); samdir.mkdir()
    genome = []
    for k, fn in enumerate(sorted(fast5_files)[:2]):
        os.symlink(fn, f5dir / os.path.basename(fn))
        read = fast5.read_fast5_arrays(fn)
        g, sam = _synthetic_alignment(rng, read.bases.tobytes().decode())
        genome.append(">chr%d\n%s\n" % (k, g))
        (samdir / (os.path.basename(fn).split(".")[0] + ".sam")).write_text("@HD\tVN:1.0\n" + sam.replace("chr1", "chr%d" % k) + "\n")
    (tmp_path / "ref.fasta").write_text("".join(genome))
    out, mdl = str(tmp_path / "out") + "/", str(tmp_path / "model")
    cmd = [sys.executable, os.path.join(root, "NanoReviser_train.py"), "-d", str(f5dir) + "/", "-o", out, "-r", str(tmp_path / "ref.fasta"),
           "-S", "tiny", "-M", mdl, "-t", str(tmp_path / "tmp") + "/", "-e", "2", "-b", "256", "-w", "11", "--validation_split", "0.05",
           "--sam_dir", str(samdir)]
    p = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert "model 2 completed" in p.stdout
    for tag, nc in (("model1", 6), ("model2", 5)):
        w = weights.load_model_weights(os.path.join(mdl, "tiny", "tiny_win11_2ep_%s.h5" % tag))
        assert w.window == 11 and w.n_class == nc
        assert os.path.exists(os.path.join(mdl, "tiny", "training_model", "train_tiny_win11_2ep_%s.h5" % tag))
        rows = open(out + "tiny_win11_2ep_%s_hisroty.csv" % tag).read().strip().split("\n")
        assert rows[0].split(",")[:5] == ["loss", "final_out_loss", "l2_loss1_loss", "final_out_acc", "l2_loss1_acc"] and len(rows) == 3
        assert float(rows[2].split(",")[0]) < float(rows[1].split(",")[0])                   # the loss goes down
        import json
        assert json.load(open(out + "tiny_win11_2ep_%s_parameters.json" % tag))["epochs"] == 2
    assert len([f for f in os.listdir(os.path.join(mdl, "tiny", "training_input")) if f.endswith(".npz")]) == 2
