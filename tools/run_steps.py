"""A few eager wide learn steps (for ncu captures; not a bench)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gold_b200 import _capi as G, synth  # noqa: E402

dims, B, N = synth.WIDE, 8192, 1 << 17
h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B, replay_capacity=N,
             precision=G.PREC_BF16)
h.set_params(G.NET_ONLINE, synth.glorot_params(dims, 1)); h.set_params(G.NET_TARGET, synth.glorot_params(dims, 2))
h.replay_push(*synth.replay(dims, N))
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    bd = h.step_breakdown(100 + i)
h.sync()
print({k: round(v * 1e3, 2) for k, v in bd})
h.close()
