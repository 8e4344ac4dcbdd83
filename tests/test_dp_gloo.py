"""world_size-2 gloo test (CPU) of the data-parallel decomposition: each rank computes the
oracle gradient of its contiguous half of the global batch, rescales it to the global mean,
the ranks all-reduce(sum), and the result must equal the single-process full-batch step.
This is the host-side contract goldq_learn implements on the GPUs with NCCL."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    from gold_b200 import dp, synth
    from oracle import binding as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dims, Bg, N = synth.CARTPOLE, 64, 512
    P = synth.nparams(dims)
    th = synth.glorot_params(dims, 1); tt = synth.glorot_params(dims, 2)
    data = synth.replay(dims, N)
    gidx = synth.sample_indices(N, Bg, 0)
    mine = dp.shard_indices(gidx, rank, world)
    s, a, r, s2, d = O.gather(dims[0], *data, mine)
    w = th.astype(np.float64)
    out = O.learn(dims, len(mine), w, tt.astype(np.float64), np.zeros(P), np.zeros(P), 0, s.astype(np.float64), a,
                  r.astype(np.float64), s2.astype(np.float64), d, dtype=np.float64)
    g = torch.tensor(dp.scale_local_grads(out["grads"], len(mine), Bg))
    loss = torch.tensor([out["loss"] * len(mine) / Bg])
    dist.all_reduce(g); dist.all_reduce(loss)
    # identical Adam on every rank
    w2 = th.copy(); m = np.zeros(P, np.float32); v = np.zeros(P, np.float32)
    O.adam(w2, g.numpy().astype(np.float32), m, v, 1)
    if rank == 0:
        q.put((g.numpy(), float(loss[0]), w2, mine.tolist()))
    else:
        q.put((None, None, w2, mine.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_decomposition_matches_full_batch_step():
    from gold_b200 import synth
    from oracle import binding as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = next(r[0] for r in res if r[0] is not None)
    loss = next(r[1] for r in res if r[1] is not None)
    np.testing.assert_array_equal(res[0][2], res[1][2])            # every rank ends with the same weights
    shards = sorted(res, key=lambda r: r[3][0] if False else 0)
    dims, Bg, N = synth.CARTPOLE, 64, 512
    P = synth.nparams(dims)
    th = synth.glorot_params(dims, 1); tt = synth.glorot_params(dims, 2)
    data = synth.replay(dims, N)
    gidx = synth.sample_indices(N, Bg, 0)
    assert sorted(res[0][3] + res[1][3]) == sorted(gidx.tolist())  # the shards partition the global batch
    s, a, r, s2, d = O.gather(dims[0], *data, gidx)
    w = th.astype(np.float64)
    full = O.learn(dims, Bg, w, tt.astype(np.float64), np.zeros(P), np.zeros(P), 0, s.astype(np.float64), a,
                   r.astype(np.float64), s2.astype(np.float64), d, dtype=np.float64)
    np.testing.assert_allclose(g, full["grads"], rtol=1e-12, atol=1e-15)
    assert abs(loss - full["loss"]) <= 1e-12 * full["loss"]
    np.testing.assert_allclose(res[0][2], w, rtol=1e-5, atol=1e-6)
