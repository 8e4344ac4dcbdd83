"""2-GPU data-parallel parity: two ranks, each training half of a global batch through goldq_learn
with a gradient all-reduce (ncclAllReduce, or on the wide path the fused peer-memory kernel),
must end with the parameters of the single-GPU full-batch step.  Skipped on boxes with fewer
than 2 GPUs (run with `gpurun --gpus 2`)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, kind, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from gold_b200 import _capi as G, dp, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dims, Bg, N, kw = _cfg(kind)
    if kind == "wide_nccl":
        os.environ["GOLDQ_NO_P2P"] = "1"
    if kind == "wide_capped":       # pretend the device keeps only 8 CTAs resident: the group must keep NCCL
        os.environ["GOLDQ_P2P_RESIDENT_CAP"] = "8"
    th = synth.glorot_params(dims, 1); tt = synth.glorot_params(dims, 2)
    data = synth.replay(dims, N)
    h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=Bg // world,
                 global_batch=Bg, replay_capacity=N, device=rank, **kw)
    h.set_params(G.NET_ONLINE, th); h.set_params(G.NET_TARGET, tt)
    h.replay_push(*data)
    dp.init_group(h, dist)
    mode, note = h.dp_mode(), h.last_error()
    w_after1 = loss_after1 = None
    for step in range(3):
        gidx = synth.sample_indices(N, Bg, step)
        h.learn(dp.shard_indices(gidx, rank, world))
        if step == 0:
            w_after1, loss_after1 = h.get_params(G.NET_ONLINE), float(h.last_loss()[0])
    w_after3, loss_after3 = h.get_params(G.NET_ONLINE), float(h.last_loss()[0])
    # goldq_learn_many (graph replay up to the all-reduce) == the same steps issued one by one
    h2 = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=Bg // world,
                  global_batch=Bg, replay_capacity=N, device=rank, **kw)
    h2.set_params(G.NET_ONLINE, w_after3); h2.set_params(G.NET_TARGET, tt)
    m, v, it = h.get_adam_state(); h2.set_adam_state(m, v, it)
    h2.replay_push(*data)
    dp.init_group(h2, dist)
    h.learn_many(21, 777 + 1000 * rank)            # graphs of 16 + 4 + 1 steps
    for i in range(21):
        h2.learn(None, 777 + 1000 * rank + i)
    same = bool(np.array_equal(h.get_params(G.NET_ONLINE), h2.get_params(G.NET_ONLINE)))
    h2.close()
    q.put((rank, w_after3, loss_after3, same, mode, note, w_after1, loss_after1))
    h.close()
    dist.barrier()
    dist.destroy_process_group()


def _cfg(kind):
    from gold_b200 import _capi as G, synth
    if kind == "narrow":
        return synth.CARTPOLE, 512, 4096, {}
    if kind == "generic":
        return (5, 17, 33, 3), 128, 1024, dict(path=G.PATH_GENERIC)
    if kind == "wide_headline":     # BASELINE configs[3]: 128->512->512->18, global batch 65536
        return synth.WIDE, 65536, 1 << 17, dict(precision=G.PREC_BF16)
    if kind == "wide_h1024":        # update grid of ~2350 CTAs: not co-resident -> the group keeps NCCL
        return (128, 1024, 1024, 18), 1024, 4096, dict(precision=G.PREC_BF16)
    return (64, 256, 128, 6), 512, 2048, dict(precision=G.PREC_BF16)      # wide, wide_nccl, wide_capped


_P2P_KINDS = ("wide", "wide_headline")


def _run_group(kind, world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("kind", ["narrow", "generic", "wide", "wide_nccl", "wide_capped", "wide_h1024", "wide_headline"])
def test_dp_equals_single_gpu(kind, world):
    from gold_b200 import _capi as G, synth
    if world > 2 and kind not in _P2P_KINDS:
        pytest.skip("NCCL modes are covered at world 2")
    res = _run_group(kind, world)
    # the wide path sums gradients with the fused peer-memory kernel (mode 2) unless told otherwise,
    # or unless the update grid cannot be co-resident (then goldq_dp_init says why and keeps NCCL)
    assert all(r[4] == (2 if kind in _P2P_KINDS else 1) for r in res), res[0][5]
    if kind in ("wide_capped", "wide_h1024"):
        assert "not co-resident" in res[0][5], res[0][5]
    for r in res[1:]:
        np.testing.assert_array_equal(res[0][1], r[1])              # ranks stay bit-identical
        assert res[0][2] == r[2]
    assert all(r[3] for r in res), "learn_many differs from single steps under data parallelism"
    dims, Bg, N, kw = _cfg(kind)
    th = synth.glorot_params(dims, 1); tt = synth.glorot_params(dims, 2)
    h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=Bg,
                 replay_capacity=N, **kw)
    h.set_params(G.NET_ONLINE, th); h.set_params(G.NET_TARGET, tt)
    h.replay_push(*synth.replay(dims, N))
    w_first = l_first = None
    for step in range(3):
        h.learn(synth.sample_indices(N, Bg, step))
        if step == 0:
            w_first, l_first = h.get_params(G.NET_ONLINE), float(h.last_loss()[0])
    w1 = h.get_params(G.NET_ONLINE); l1 = float(h.last_loss()[0]); h.close()
    # same arithmetic, different summation grouping (per-rank partial sums): fp32 noise only,
    # amplified on near-zero gradients by Adam's sign-like early steps
    assert abs(res[0][2] - l1) <= (1e-3 if kind.startswith("wide") else 1e-5) * abs(l1)
    dw = np.abs(res[0][1] - w1)
    assert dw.max() <= 2.5e-3 and np.mean(dw > 1e-6) < 0.05

    if kind == "wide_headline":
        # ... and against the fp64 oracle after the first step, at the wide path's stated tolerance
        # (tests/test_gpu_wide.py): loss 1e-2, parameters within 1e-2 of the weight scale and inside
        # Adam's hard bound
        from oracle import binding as O
        O.set_threads(O.max_threads())
        data = synth.replay(dims, N)
        idx = synth.sample_indices(N, Bg, 0)
        sb, ab, rb, s2b, db = O.gather(dims[0], *data, idx)
        P = synth.nparams(dims)
        w64 = th.astype(np.float64); m64 = np.zeros(P); v64 = np.zeros(P)
        out64 = O.learn(dims, Bg, w64, tt.astype(np.float64), m64, v64, 0, sb.astype(np.float64), ab,
                        rb.astype(np.float64), s2b.astype(np.float64), db, dtype=np.float64)
        O.set_threads(1)
        for w_gpu, l_gpu in ((res[0][6], res[0][7]), (w_first, l_first)):
            assert abs(l_gpu - out64["loss"]) <= 1e-2 * abs(out64["loss"]), (l_gpu, out64["loss"])
            dw64 = np.abs(w_gpu - w64)
            assert dw64.max() <= 2 * 5e-4 * (1 + np.abs(out64["grads"]).max()) + 1e-7
            assert dw64.max() <= 1e-2 * np.abs(w64).max()
