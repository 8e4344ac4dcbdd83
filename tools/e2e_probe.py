"""Where does the host-buffer (e2e) step go?  (diagnostic)  Times, on the wide workload: the packed push alone, the learn
step alone (host indices), and both as bench.py's e2e leg issues them."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gold_b200 import _capi as G, synth  # noqa: E402

dims, B, N = synth.WIDE, 8192, 1 << 17
h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B, replay_capacity=N,
             precision=G.PREC_BF16)
h.set_params(G.NET_ONLINE, synth.glorot_params(dims, 1)); h.set_params(G.NET_TARGET, synth.glorot_params(dims, 2))
h.replay_push(*synth.replay(dims, N))
h.learn_prepare()
nbytes = h.packed_bytes(B)
L = G.lib()


def pinned(n, dt):
    p = L.goldq_host_alloc(int(n) * np.dtype(dt).itemsize)
    return np.frombuffer((C.c_uint8 * (int(n) * np.dtype(dt).itemsize)).from_address(p), dtype=dt), p


bufs = []
for k in range(2):
    arr, ptr = pinned(nbytes, np.uint8)
    h.pack(*synth.replay(dims, B, seed=50 + k), out=arr)
    bufs.append(ptr)
idx, idx_ptr = pinned(B, np.int32)
idx[:] = synth.sample_indices(N, B, 3)
K = 100


def timed(name, fn):
    for i in range(5):
        fn(i)
    h.sync()
    t0 = time.perf_counter()
    for i in range(K):
        fn(i)
    h.sync()
    dt = (time.perf_counter() - t0) / K * 1e6
    print(f"{name:58s} {dt:8.1f} us/step", flush=True)


timed("push only (one 8.5 MB packed H2D + scatter), no sync", lambda i: h.replay_push_packed(B, bufs[i & 1]))
timed("push + wait for the copy each time", lambda i: (h.replay_push_packed(B, bufs[i & 1]), h.replay_push_wait()))
timed("learn only (host indices, graph), no sync", lambda i: h.learn_raw(idx_ptr))
timed("learn + last_loss (sync)", lambda i: (h.learn_raw(idx_ptr), h.last_loss()))
timed("learn, push next, last_loss (bench.py's e2e step)", lambda i: (h.learn_raw(idx_ptr), h.replay_push_packed(B, bufs[(i + 1) & 1]), h.last_loss()))
timed("push next, learn, last_loss", lambda i: (h.replay_push_packed(B, bufs[(i + 1) & 1]), h.learn_raw(idx_ptr), h.last_loss()))
timed("learn, push next, NO per-step sync", lambda i: (h.learn_raw(idx_ptr), h.replay_push_packed(B, bufs[(i + 1) & 1])))
timed("learn (device sampler), push next, last_loss", lambda i: (h.learn(None, seed=i), h.replay_push_packed(B, bufs[(i + 1) & 1]), h.last_loss()))
timed("learn, push next, sync()", lambda i: (h.learn_raw(idx_ptr), h.replay_push_packed(B, bufs[(i + 1) & 1]), h.sync()))
h.replay_stage_packed(B, bufs[0])
timed("learn, commit, stage next-but-one, last_loss", lambda i: (h.learn_raw(idx_ptr), h.replay_commit(), h.replay_stage_packed(B, bufs[(i + 1) & 1]), h.last_loss()))
h.close()
