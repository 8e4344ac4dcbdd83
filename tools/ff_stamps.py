"""In-kernel timeline of the fused forward on CTA pairs (diagnostic).  Run with GOLDQ_FF_STAMPS=1.
Prints, per stamp, the median / max over CTAs of (stamp - stamp[1]) in SM cycles."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GOLDQ_FF_STAMPS"] = "1"
from gold_b200 import _capi as G, synth  # noqa: E402

NAMES = [""] * 96
NAMES[0], NAMES[1], NAMES[2], NAMES[9] = "entry", "after cluster_sync+pdl_wait", "MMA: x_full seen", "MMA: all issued"
for p_, nm in enumerate(["L1 half0", "L1 half1", "L2 half0", "L2 half1"]):
    NAMES[10 + p_] = f"EPI: {nm} accumulator ready"
    NAMES[14 + p_] = f"EPI: {nm} written, arrived"
NAMES[18], NAMES[19] = "EPI: Q accumulator ready", "EPI: Q written, H stores read out"
NAMES[20], NAMES[21] = "EPI: Q staged", "EPI: Q copied out"
for blk in range(24):
    NAMES[32 + 2 * blk], NAMES[33 + 2 * blk] = f"MMA: stage {blk} operands ready", f"MMA: stage {blk} issued"
for i in range(16):
    NAMES[80 + i] = f"TMA: load {i} slot free, issuing"
dims, B, N = synth.WIDE, 8192, 1 << 16
h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B, replay_capacity=N,
             precision=G.PREC_BF16)
h.set_params(G.NET_ONLINE, synth.glorot_params(dims, 1)); h.set_params(G.NET_TARGET, synth.glorot_params(dims, 2))
h.replay_push(*synth.replay(dims, N))
for i in range(5):
    h.learn(None, i)
h.sync()
for i in range(3):
    bd = h.step_breakdown(100 + i)
print("eager breakdown (us):", {k: round(v * 1e3, 2) for k, v in bd})
st = h.debug_fetch(G.DBG_FF_STAMPS).astype(np.int64)
nx = B // 128
g0, g1 = st[:, 30], st[:, 31]
print(f"globaltimer: first CTA entry -> last CTA exit {(g1.max() - g0.min()) / 1e3:.2f} us; entry skew {(g0.max() - g0.min()) / 1e3:.2f} us; "
      f"exit skew {(g1.max() - g1.min()) / 1e3:.2f} us; median CTA lifetime {np.median(g1 - g0) / 1e3:.2f} us")
st = st.reshape(2, nx, 96)
for label, sel in (("ONLINE net, leader CTAs", st[0:1, 0::2]), ("TARGET net, leader CTAs", st[1:2, 0::2])):
    rel = sel - sel[..., 1:2]
    print(label)
    order = np.argsort(np.median(rel.reshape(-1, 96), axis=0))
    for i in order:
        col = rel[..., i].ravel()
        if (sel[..., i] == 0).all():
            continue
        print(f"  {NAMES[i]:42s} median {np.median(col):8.0f}  min {col.min():8.0f}  max {col.max():8.0f}")
h.close()
