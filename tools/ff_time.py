"""Forward kernel alone (20 back-to-back launches) and the graph-replayed step, for whatever GOLDQ_* switches
the environment carries (diagnostic; one line of output, meant for A/B runs in one gpurun call)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gold_b200 import _capi as G, synth  # noqa: E402

dims, B, N = synth.WIDE, 8192, 1 << 17
h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B, replay_capacity=N,
             precision=G.PREC_BF16)
h.set_params(G.NET_ONLINE, synth.glorot_params(dims, 1)); h.set_params(G.NET_TARGET, synth.glorot_params(dims, 2))
h.replay_push(*synth.replay(dims, N))
for i in range(3):
    h.learn(None, i)
h.sync()
ff = [h.time_forward_us(20) for _ in range(5)]
h.learn_prepare()
h.learn_many(64, 7)
h.sync()
steps = []
for r in range(5):
    t0 = time.perf_counter()
    h.learn_many(256, 100 + r)
    h.sync()
    steps.append((time.perf_counter() - t0) / 256 * 1e6)
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("GOLDQ_"))
print(f"[{tag or 'default'}] forward us/launch min {min(ff):.2f} median {sorted(ff)[2]:.2f} | step us (host clock, 256-step replays) "
      f"min {min(steps):.2f} median {sorted(steps)[2]:.2f} | loss {float(h.last_loss()[0]):.6f}")
h.close()
