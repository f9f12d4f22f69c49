"""Training throughput of the GPU path (windows per second through forward + backward + Adam) at the reference's batch size, beside
the same step in torch fp32 autograd on the host cores (the graph of tests/test_train_gpu.py).  Not a bench.py metric: the
north-star metric is inference; this is the context number for SURVEY.md section 8(f) rank 4.
usage (GPU box): python tools/bench_train.py [batch] [window] > gpurun_out/<tag>_train.json"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from nanoreviser_b200 import train  # noqa: E402
import test_train_gpu as TT  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
T = int(sys.argv[2]) if len(sys.argv) > 2 else 13
rng = np.random.default_rng(0)
S, X, y, _ = TT._inputs(rng, B, T, 6)
X[..., 4:6] /= 100.0
tm = train.TrainModel(window=T, n_class=6, seed=1)
GRAPH = os.environ.get("NRV_TRAIN_GRAPH", "1") != "0"
for _ in range(4):
    tm.train_on_batch(S, X, y, graph=GRAPH)
torch.cuda.synchronize()
steps = 20
t0 = time.perf_counter()
for _ in range(steps):
    tm.train_on_batch(S, X, y, graph=GRAPH)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / steps
# host baseline: the same graph in torch fp32 autograd + torch.optim.Adam on the CPU
P = {k: v.detach().cpu().float().clone().requires_grad_(True) for k, v in tm.p.items()}
opt = torch.optim.Adam(list(P.values()), lr=1e-3, eps=1e-7)
Sc, Xc, yc = torch.tensor(S), torch.tensor(X), torch.tensor(y)
mask = torch.ones(B, T, 50, 8)
def cpu_step():
    opt.zero_grad()
    loss, *_ = TT._ref_graph(torch, P, Sc, Xc, yc, mask, None, 6)
    loss.backward()
    opt.step()
cpu_step()
t0 = time.perf_counter()
for _ in range(3):
    cpu_step()
dc = (time.perf_counter() - t0) / 3
print(json.dumps({"metric": "training_windows_per_sec", "batch": B, "window": T, "value": B / dt, "ms_per_step": dt * 1e3,
                  "cuda_graph": GRAPH, "operator_calls_per_step": 589, "cpu_torch_fp32_autograd": {"value": B / dc, "ms_per_step": dc * 1e3,
                  "threads": torch.get_num_threads()}, "note": "forward + backward + Adam on one batch, inputs uploaded every step"}))
