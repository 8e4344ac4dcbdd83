"""Which launches of the wide step sit on its critical path (diagnostic; results with GOLDQ_SKIP are garbage).
Step time of the 16-step graphs with one launch left out at a time."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gold_b200 import _capi as G, synth  # noqa: E402

NAMES = {0: "(nothing skipped)", 1: "gather", 2: "forward", 4: "TD/dZ2", 8: "dW3", 16: "dW2", 32: "dX", 64: "db1 colsum",
         128: "dW1", 256: "update", 8 + 16: "dW3+dW2", 8 + 64: "dW3+db1", 4 + 8 + 16 + 32 + 64 + 128: "whole backward",
         2 + 4 + 8 + 16 + 32 + 64 + 128 + 256: "all but gather", 511: "everything"}
dims, B, N = synth.WIDE, 8192, 1 << 17
data = synth.replay(dims, N)
th, tt = synth.glorot_params(dims, 1), synth.glorot_params(dims, 2)
base = None
for mask in NAMES:
    os.environ["GOLDQ_SKIP"] = str(mask)
    h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B, replay_capacity=N,
                 precision=G.PREC_BF16)
    h.set_params(G.NET_ONLINE, th); h.set_params(G.NET_TARGET, tt)
    h.replay_push(*data)
    h.learn_prepare()
    h.learn_many(64, 1); h.sync()
    ts = []
    for rep in range(5):
        h.timer_start(); h.learn_many(96, 1000 + 100 * rep); ts.append(h.timer_stop() / 96 * 1e3)
    t = float(np.median(ts))
    base = t if base is None else base
    print(f"skip {NAMES[mask]:20s} {t:7.2f} us/step   delta {base - t:6.2f}", flush=True)
    h.close()
