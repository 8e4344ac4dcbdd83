"""Small run of every path for compute-sanitizer (memcheck / racecheck)."""
import sys
import numpy as np
sys.path.insert(0, '.')
from gold_b200 import _capi as G, synth

def run(dims, B, N, **kw):
    R = kw.get('n_replicas', 1)
    h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B, replay_capacity=N, **kw)
    h.set_params(G.NET_ONLINE, np.concatenate([synth.glorot_params(dims, 1 + r) for r in range(R)]))
    h.set_params(G.NET_TARGET, np.concatenate([synth.glorot_params(dims, 9 + r) for r in range(R)]))
    parts = [synth.replay(dims, N, seed=3 + r) for r in range(R)]
    h.replay_push(*[np.concatenate([p[k] for p in parts]) for k in range(5)])
    h.learn(np.stack([synth.sample_indices(N, B, r) for r in range(R)]))
    h.learn(None, seed=5)
    h.learn_many(17, seed=9)
    h.sync_target()
    h.predict(G.NET_ONLINE, parts[0][0][:7])
    h.greedy_action(parts[0][0][:7])
    print(dims, kw, 'loss', h.last_loss()[:2], flush=True)
    h.close()

run(synth.CARTPOLE, 64, 512)
run(synth.CARTPOLE, 20, 128, n_replicas=3)
run((5, 17, 33, 3), 37, 256, path=G.PATH_GENERIC)
run(synth.WIDE, 256, 1024, precision=G.PREC_BF16)
run((64, 256, 128, 6), 200, 1024, precision=G.PREC_BF16)
print('done')
