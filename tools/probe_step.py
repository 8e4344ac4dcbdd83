"""Diagnostic: GPU latency of one learn step vs host issue rate (not a bench)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from gold_b200 import _capi as G, synth

def probe(name, **kw):
    dims = (kw['obs_dim'], kw['hidden1'], kw['hidden2'], kw['n_actions'])
    R = kw.get('n_replicas', 1)
    N = kw['replay_capacity']
    h = G.Handle(**kw)
    h.set_params(G.NET_ONLINE, np.concatenate([synth.glorot_params(dims, 1 + r) for r in range(R)]))
    h.set_params(G.NET_TARGET, np.concatenate([synth.glorot_params(dims, 50 + r) for r in range(R)]))
    parts = [synth.replay(dims, N, seed=9 + r) for r in range(R)]
    h.replay_push(*[np.concatenate([p[k] for p in parts]) for k in range(5)])
    for i in range(20):
        h.learn(None, seed=i)
    h.sync()
    single = []
    for i in range(20):
        h.sync(); h.timer_start(); h.learn(None, seed=100 + i); single.append(h.timer_stop())
    h.sync()
    t0 = time.perf_counter(); h.timer_start()
    K = 320
    for i in range(K):
        h.learn(None, seed=200 + i)
    t_issue = time.perf_counter() - t0
    ms = h.timer_stop()
    h.sync()
    for i in range(3):
        h.learn_many(32, 5000 + 32 * i)
    h.sync()
    t0 = time.perf_counter(); h.timer_start()
    for i in range(K // 32):
        h.learn_many(32, 9000 + 32 * i)
    t_issue_g = time.perf_counter() - t0
    ms_g = h.timer_stop()
    print(f"{name}: graph replay (learn_many(32)) {ms_g/K*1e3:.1f} us/step; host issue {t_issue_g/K*1e6:.2f} us/step")
    print(f"{name}: single-step GPU latency median {np.median(single)*1e3:.1f} us (min {min(single)*1e3:.1f}); "
          f"back-to-back {ms/K*1e3:.1f} us/step; host issue {t_issue/K*1e6:.1f} us/step; launches/step "
          f"{h.launch_count()/(K+40):.1f}", flush=True)
    h.close()

if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if which in ('all', 'cartpole'):
        probe('cartpole B=4096', obs_dim=4, hidden1=24, hidden2=24, n_actions=2, batch_size=4096, replay_capacity=65536)
    if which in ('all', 'replicas'):
        probe('replicas 128xB=20', obs_dim=4, hidden1=24, hidden2=24, n_actions=2, batch_size=20, replay_capacity=4096, n_replicas=128)
    if which in ('all', 'wide'):
        probe('wide bf16 B=8192', obs_dim=128, hidden1=512, hidden2=512, n_actions=18, batch_size=8192,
              replay_capacity=int(__import__('os').environ.get('PROBE_ROWS', 65536)), precision=G.PREC_BF16)
    if which in ('all', 'widefp32'):
        probe('wide fp32 generic B=8192', obs_dim=128, hidden1=512, hidden2=512, n_actions=18, batch_size=8192,
              replay_capacity=65536)
