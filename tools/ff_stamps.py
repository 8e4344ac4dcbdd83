"""In-kernel timeline of the fused forward on CTA pairs (diagnostic).  Run with GOLDQ_FF_STAMPS=1.
Prints, per stamp, the median / max over CTAs of (stamp - stamp[1]) in SM cycles."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["GOLDQ_FF_STAMPS"] = "1"
from gold_b200 import _capi as G, synth  # noqa: E402

NAMES = ["entry", "after cluster_sync+pdl_wait", "MMA: x_full seen", "MMA: L1 issued", "MMA: L2 start (h_lo)",
         "MMA: L2 kb4 gate passed", "MMA: L2 issued", "MMA: L3 start", "MMA: L3 gate passed", "MMA: L3 issued",
         "EPI: acc_full[0] L1", "EPI: H1 complete", "EPI: acc_full[0] L2", "EPI: acc_full[1] L2 (all L2 MMAs done)",
         "EPI: H2 complete", "EPI: acc_full[0] L3", "EPI: Q written", "MMA: reached L2 kb4 gate", "MMA: reached L3 gate",
         "EPI: half0 held, waiting L2 end", "EPI: H2 half0 stored", "MMA: h_lo(L2) seen", "MMA: h_ready(L2) seen",
         "MMA: h_lo(L3) seen", "MMA: h_ready(L3) seen", "EPI: L1 half0 chunks done", "EPI: L1 h0 proxy fence done",
         "EPI: L1 h0 bar.sync done", "EPI: L1 h0 h_lo arrived", "", "", ""]
dims, B, N = synth.WIDE, 8192, 1 << 16
h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B, replay_capacity=N,
             precision=G.PREC_BF16)
h.set_params(G.NET_ONLINE, synth.glorot_params(dims, 1)); h.set_params(G.NET_TARGET, synth.glorot_params(dims, 2))
h.replay_push(*synth.replay(dims, N))
for i in range(5):
    h.learn(None, i)
h.sync()
st = h.debug_fetch(G.DBG_FF_STAMPS).astype(np.int64)
nx = B // 128
st = st.reshape(2, nx, 32)
for label, sel in (("ONLINE net, leader CTAs", st[0:1, 0::2]), ("ONLINE net, peer CTAs", st[0:1, 1::2]),
                   ("TARGET net, leader CTAs", st[1:2, 0::2])):
    rel = sel - sel[..., 1:2]
    print(label)
    order = np.argsort(np.median(rel.reshape(-1, 32), axis=0))
    for i in order:
        col = rel[..., i].ravel()
        if (sel[..., i] == 0).all():
            continue
        print(f"  {NAMES[i]:42s} median {np.median(col):8.0f}  min {col.min():8.0f}  max {col.max():8.0f}")
h.close()
