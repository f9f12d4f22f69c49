#!/usr/bin/env python
"""cfg4 (BASELINE.json configs[3]): Bi-LSTM recurrence / Conv1D batch sweep, 1..65536 windows per launch, ecoli and
human weights.  One synthetic read of B + W bases gives exactly B windows; per-stage device times come from the
library's CUDA-event stage timers (K2 = cnn, K3 = read_rnn1..heads).  Prints a markdown table.

  python tools/sweep_batch.py > profiles/r01_batch_sweep.md        (on the GPU box)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nanoreviser_b200 import engine, synth, weights  # noqa: E402

K3 = ("lstm0", "proj1", "rec1", "proj2", "rec2", "proj3", "rec3", "heads_gemm", "heads")
MAC_K3 = 2 * (30_976 + 540_672 + 3_604_480 + 1_802_240 + 228_544)      # both models, per window
MAC_K2 = 2 * 36_400                                                     # both models, per base


def main():
    print("# cfg4: windows-per-launch sweep (device time per launch, CUDA events; 3 warm-up + 5 timed)\n")
    print("| species | windows | K2 cnn us | K3 model us | K3 us/window | K3 TFLOP/s (algorithmic) | K3 windows/s |")
    print("|---|---|---|---|---|---|---|")
    for sp in ("ecoli", "human"):
        m1, m2 = weights.load_species(sp, os.path.join(ROOT, "model"))
        with engine.Reviser(m1, m2, device=0) as rv:
            W = rv.window
            for e in range(0, 17):
                B = 1 << e
                b = synth.make_batch([B + W], seed=e)
                for _ in range(3):
                    rv.revise_batch(b)
                rv.set_stage_timing(True)
                n = 5
                for _ in range(n):
                    rv.revise_batch(b)
                rv.synchronize()
                ms = rv.stage_ms()
                rv.set_stage_timing(False)
                k2 = ms.get("cnn", 0.0) / n * 1e3
                k3 = sum(ms.get(k, 0.0) for k in K3) / n * 1e3
                print("| %s | %d | %.1f | %.1f | %.3f | %.2f | %.3g |" % (
                    sp, B, k2, k3, k3 / B, 2.0 * MAC_K3 * B / (k3 * 1e-6) / 1e12, B / (k3 * 1e-6)))
    print("\nLatency floor: a launch of up to 128 windows per direction is one tile per kernel (9 kernels per model); "
          "throughput saturates once every SM has a tile (>= 148 x 128 = 18,944 windows per model chunk).")


if __name__ == "__main__":
    main()
