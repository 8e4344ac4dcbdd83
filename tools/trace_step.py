"""Timeline of graph-replayed learn steps on the wide path (diagnostic; there is no nsys in this image).

GOLDQ_TRACE=1 makes thread 0 of every block of the step's kernels log {kernel, block, globaltimer at entry / exit}
(BlockTrace, kernels_wide.cu).  This tool replays a few 16-step graphs, fetches the log, cuts it into kernel
instances (same kernel id, blocks overlapping or close in time) and prints, for the median step of the replay, when
each kernel's first block started, when its last block ended, and how busy the SMs were in between.

    python tools/trace_step.py [steps]        (set GOLDQ_* switches in the environment as usual)
"""
import os
import sys

import numpy as np

os.environ.setdefault("GOLDQ_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("GOLDQ_LIB", os.path.join(ROOT, "gold_b200", "lib", "libgoldq_trace.so"))   # make -C gold_b200/csrc trace
sys.path.insert(0, ROOT)
from gold_b200 import _capi as G, synth  # noqa: E402

NAMES = {1: "gather", 2: "forward", 3: "TD/dZ2", 4: "dW3", 5: "dW2", 6: "dX", 7: "colsum", 8: "dW1", 9: "update", 10: "other"}
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dims, B, N = synth.WIDE, int(os.environ.get("PROBE_B", 8192)), 1 << 17
h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B, replay_capacity=N,
             precision=G.PREC_BF16)
h.set_params(G.NET_ONLINE, synth.glorot_params(dims, 1)); h.set_params(G.NET_TARGET, synth.glorot_params(dims, 2))
h.replay_push(*synth.replay(dims, N))
h.learn_prepare()
h.learn_many(32, 1)
h.sync()
h.debug_fetch(G.DBG_TRACE)                 # drop the warm-up's records
h.timer_start()
h.learn_many(steps, 100)
ms = h.timer_stop()
raw = h.debug_fetch(G.DBG_TRACE)
n = min(int(raw[0]), (len(raw) - 4) // 12)       # (records past the buffer's capacity were dropped)
rec = raw[4:4 + 12 * n].reshape(n, 12)
marks = rec[:, 4:12].astype(np.int64)
t0, t1 = rec[:, 0].astype(np.int64), rec[:, 1].astype(np.int64)
kid = (rec[:, 2] & np.uint64(0xffffffff)).astype(np.int64)
print(f"{steps} steps, {ms * 1e3 / steps:.2f} us/step with tracing on, {n} block records")
base = t0.min()
t0 -= base; t1 -= base

# kernel instances: per id, sort blocks by start; a new instance starts when a block starts after every earlier
# block of the instance has ended by more than `gap` ns
inst = []
for k in np.unique(kid):
    sel = np.where(kid == k)[0]
    order = sel[np.argsort(t0[sel])]
    cur_s, cur_e, cnt, busy = None, None, 0, 0
    for i in order:
        if cur_s is None or t0[i] > cur_e + 1500:
            if cur_s is not None:
                inst.append((cur_s, cur_e, int(k), cnt, busy))
            cur_s, cur_e, cnt, busy = t0[i], t1[i], 0, 0
        cur_e = max(cur_e, t1[i]); cnt += 1; busy += t1[i] - t0[i]
    inst.append((cur_s, cur_e, int(k), cnt, busy))
inst.sort()
# steps are delimited by the update kernel
ups = [x for x in inst if x[2] == 9]
print(f"{len(ups)} update instances; step period (update end to update end), us:",
      np.round(np.diff([u[1] for u in ups]) / 1e3, 2)[:40])
if len(ups) >= 6:
    per = np.diff([u[1] for u in ups])
    mid = int(np.argsort(per)[len(per) // 2]) + 1          # a step of median length (not a graph boundary)
    lo, hi = ups[mid - 1][1], ups[mid][1]
    print(f"\nstep {mid}: window = end of update {mid - 1} .. end of update {mid} = {(hi - lo) / 1e3:.2f} us   (times in us from the window start)")
    print(f"{'kernel':10s} {'first block in':>14s} {'last block out':>15s} {'span':>7s} {'blocks':>7s} {'mean block':>11s}")
    for s, e, k, c, busy in inst:
        if e > lo - 20000 and s < hi + 5000 and (s >= lo - 12000):
            print(f"{NAMES.get(k, k):10s} {(s - lo) / 1e3:14.2f} {(e - lo) / 1e3:15.2f} {(e - s) / 1e3:7.2f} {c:7d} {busy / c / 1e3:11.2f}")
    # in-kernel marks (dX): relative to mark 0 (= after griddepcontrol.wait), blocks of the instances inside the window
    MARKS = {6: ["after pdl_wait", "MMA: first stage landed", "MMA: all issued", "EPI: accumulator ready", "EPI: tile staged",
                 "EPI: column sums done", "EPI: TMA store read out", "(EPI: column-sum loop done)"]}
    GEMM = ["after pdl_wait", "MMA: first stage landed", "MMA: all issued", "EPI: accumulator ready", "EPI: tile written"]
    MARKS.update({4: GEMM, 5: GEMM, 8: GEMM})
    MARKS[9] = ["after pdl_wait", "step scalars arrived", "slices summed (thread 0)", "Adam + stores issued"]
    MARKS[3] = ["after pdl_wait", "TD rows done (warp 0)", "dQ^T + row partials stored", "dZ2 stored (warp 0)", "block barrier passed", "(actions arrived)", "(first-max done)"]
    marks -= base
    for k, names in MARKS.items():
        sel = np.where((kid == k) & (t1 > lo) & (t1 <= hi + 3000) & (marks[:, 0] > 0))[0]
        if len(sel) == 0:
            continue
        print(f"\n{NAMES[k]}: marks of {len(sel)} blocks, us after the block passed griddepcontrol.wait (median / min / max); "
              f"wait passed at {np.median(marks[sel, 0] - lo) / 1e3:.2f} us in the window")
        for j, nm in enumerate(names):
            v = (marks[sel, j] - marks[sel, 0]) / 1e3
            v = v[marks[sel, j] > 0]
            if len(v):
                print(f"  {nm:28s} {np.median(v):7.2f} {v.min():7.2f} {v.max():7.2f}")
        v = (t1[sel] - marks[sel, 0]) / 1e3
        print(f"  {'block exit':28s} {np.median(v):7.2f} {v.min():7.2f} {v.max():7.2f}")
    # when the blocks of the critical chain start and end (deciles, us in the window): shows waves
    for k in (3, 6, 8):
        sel = np.where((kid == k) & (t1 > lo) & (t1 <= hi + 3000))[0]
        if len(sel):
            qs = [0, 0.1, 0.25, 0.5, 0.75, 0.9, 1.0]
            print(f"\n{NAMES[k]}: block start quantiles (0, .1, .25, .5, .75, .9, 1) {np.round(np.quantile((t0[sel] - lo) / 1e3, qs), 2)}")
            print(f"  block end quantiles {np.round(np.quantile((t1[sel] - lo) / 1e3, qs), 2)}")
            v = marks[sel, 0]
            v = v[v > 0]
            if len(v):
                print(f"  griddepcontrol.wait passed, quantiles {np.round(np.quantile((v - lo) / 1e3, qs), 2)}")
    # update kernel: duration by block (the block ranges depend on the lanes-per-group plan: report quantiles over all blocks
    # and the slowest tenth's block indices instead of guessing the ranges)
    blk = (rec[:, 2] >> np.uint64(32)).astype(np.int64)
    sel = np.where((kid == 9) & (t1 > lo + 5000) & (t1 <= hi + 100))[0]
    if len(sel):
        s0 = t0[sel].min()
        du, out_ = (t1[sel] - t0[sel]) / 1e3, (t1[sel] - s0) / 1e3
        print(f"\nupdate kernel, {len(sel)} blocks: duration us min / median / p90 / max = {du.min():.2f} / {np.median(du):.2f} / "
              f"{np.quantile(du, 0.9):.2f} / {du.max():.2f}; last block out {out_.max():.2f} us after the first block in")
        late = sel[np.argsort(t1[sel])[-max(1, len(sel) // 10):]]
        print(f"  blocks that finish last (index range {blk[late].min()}..{blk[late].max()}, median index {int(np.median(blk[late]))})")
h.close()
