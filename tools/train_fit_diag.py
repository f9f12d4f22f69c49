"""Diagnostic: fit on the synthetic colour task; per-epoch train / validation metrics, and the train set through the validation graph."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nanoreviser_b200 import train
import test_train_gpu as TT
rng = np.random.default_rng(5)
T, N = 11, 1536
# a learnable task shaped like the real one: every base has a signal level, the label is the centre base
_, X, _, _ = TT._inputs(rng, N, T, 6)
base = rng.integers(0, 4, (N, T))
X[..., 0] = np.array([250, 180, 100, 30])[base] / 300.0
X[..., 4:6] /= 100.0        # (raw event means ~100 saturate every gate of a freshly initialised read_rnn1)
S = (np.array([-1.0, -0.3, 0.3, 1.0])[base][..., None] + 0.1 * rng.normal(size=(N, T, 50))).astype(np.float32)
y = np.array([5, 4, 3, 2])[base[:, T // 2]]
tm = train.TrainModel(window=T, n_class=6, seed=11)
h = tm.fit([S[..., None], X, y.reshape(-1, 1)], None, validation_split=0.125, epochs=int(sys.argv[1]) if len(sys.argv) > 1 else 25, batch_size=64, verbose=1)
m = tm.forward_backward(S[:512], X[:512], y[:512], training=False)
print("train windows through the validation graph:", m)
for k in ("bn1", "bnr0", "bnr2"):
    print(k, "moving mean", tm.s[k + "_mean"][:4].tolist(), "batch", tm._buf[k + "_bm"][:4].tolist(), "moving var", tm.s[k + "_var"][:4].tolist(), "batch", tm._buf[k + "_bv"][:4].tolist())
