"""Block trace of data-parallel steps (diagnostic): what the in-kernel gradient exchange adds to the update kernel.

    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/trace_dp.py        (needs `make -C gold_b200/csrc trace`)

Every rank runs the wide configuration with 8192 rows; rank 0 prints, for one mid-graph step, the timeline of the step and
the update kernel's in-kernel marks (thread 0 of every weight block): slices summed -> exchange done -> Adam + stores."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("GOLDQ_TRACE", "1")
os.environ.setdefault("GOLDQ_LIB", os.path.join(ROOT, "gold_b200", "lib", "libgoldq_trace.so"))
sys.path.insert(0, ROOT)
from gold_b200 import _capi as G, synth, dp  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dims, B, N = synth.WIDE, 8192, 1 << 17
h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B, replay_capacity=N,
             precision=G.PREC_BF16, device=local, global_batch=B * world)
h.set_params(G.NET_ONLINE, synth.glorot_params(dims, 1)); h.set_params(G.NET_TARGET, synth.glorot_params(dims, 2))
h.replay_push(*synth.replay(dims, N, seed=1234 + rank))
dp.init_group(h, dist)
h.learn_prepare()
h.learn_many(32, 1); h.sync()
dist.barrier()
h.debug_fetch(G.DBG_TRACE)
h.timer_start()
h.learn_many(32, 100)
ms = h.timer_stop()
raw = h.debug_fetch(G.DBG_TRACE)
n = min(int(raw[0]), (len(raw) - 4) // 12)
rec = raw[4:4 + 12 * n].reshape(n, 12)
t0, t1 = rec[:, 0].astype(np.int64), rec[:, 1].astype(np.int64)
kid = (rec[:, 2] & np.uint64(0xffffffff)).astype(np.int64)
marks = rec[:, 4:12].astype(np.int64)
NAMES = {1: "gather", 2: "forward", 3: "TD/dZ2", 4: "dW3", 5: "dW2", 6: "dX", 7: "colsum", 8: "dW1", 9: "update"}
for r in range(world):
    dist.barrier()
    if r != rank:
        continue
    print(f"\n=== rank {rank} of {world} (dp mode {h.dp_mode()}): {ms * 1e3 / 32:.2f} us/step with tracing on")
    ups = np.where(kid == 9)[0]
    # update instances: cluster by time
    order = ups[np.argsort(t0[ups])]
    inst, cur = [], [order[0]]
    for i in order[1:]:
        if t0[i] > max(t1[j] for j in cur[-64:]) + 1500:
            inst.append(cur); cur = []
        cur.append(i)
    inst.append(cur)
    ends = [max(t1[i] for i in c) for c in inst]
    per = np.diff(ends)
    mid = int(np.argsort(per)[len(per) // 2]) + 1
    lo, hi = ends[mid - 1], ends[mid]
    print(f"step {mid}: {(hi - lo) / 1e3:.2f} us")
    for k in sorted(NAMES):
        sel = np.where((kid == k) & (t1 > lo) & (t1 <= hi))[0]
        if len(sel):
            print(f"  {NAMES[k]:8s} first in {(t0[sel].min() - lo) / 1e3:7.2f}  last out {(t1[sel].max() - lo) / 1e3:7.2f}  blocks {len(sel)}")
    c = np.array(inst[mid])
    c = c[marks[c, 2] > 0]
    if len(c):
        def q(v): return f"{np.median(v) / 1e3:6.2f} / {np.quantile(v, 0.9) / 1e3:6.2f} / {v.max() / 1e3:6.2f}"
        print("  update kernel, weight blocks, us (median / p90 / max):")
        print("    wait passed -> slices summed   ", q(marks[c, 2] - marks[c, 0]))
        ex = c[marks[c, 4] > 0]
        if len(ex):
            print("    slices summed -> exchange done ", q(marks[ex, 4] - marks[ex, 2]))
            print("    exchange done -> Adam + stores ", q(marks[ex, 3] - marks[ex, 4]))
        print("    block total                    ", q(t1[c] - t0[c]))
h.close()
dist.destroy_process_group()
