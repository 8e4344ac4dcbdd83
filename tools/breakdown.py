"""Diagnostic: per-launch times of one eager, serialised wide step (goldq_step_breakdown)."""
import sys
import numpy as np
sys.path.insert(0, '.')
from gold_b200 import _capi as G, synth
dims = synth.WIDE
h = G.Handle(obs_dim=128, hidden1=512, hidden2=512, n_actions=18, batch_size=8192, replay_capacity=65536, precision=G.PREC_BF16)
h.set_params(G.NET_ONLINE, synth.glorot_params(dims, 1)); h.set_params(G.NET_TARGET, synth.glorot_params(dims, 2))
h.replay_push(*synth.replay(dims, 65536))
for i in range(5): h.step_breakdown(i)
reps = [h.step_breakdown(100 + i) for i in range(20)]
ms = np.array([[t for _, t in r] for r in reps]).mean(axis=0)
for (l, _), t in zip(reps[0], ms): print(f"{t*1e3:7.1f} us  {l}")
print(f"{ms.sum()*1e3:7.1f} us  total (serialised, eager)")
