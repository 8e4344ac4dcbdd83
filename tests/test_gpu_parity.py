"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the
same weights, replay contents and sampled indices.

Bars:  sampled indices, greedy actions, Q(s), Q'(s'), TD targets — bit-exact
       (fp32 paths keep the reference's k-sequential multiply-then-add);
       loss, gradients — 1e-5 relative against the fp64 oracle twin (the
       reference's own sequential fp32 sums carry more error than that);
       post-Adam parameters — 1e-5 relative (ill-conditioned elements: see util).
"""
import numpy as np
import pytest

from gold_b200 import _capi as G
from gold_b200 import synth
from util import OracleAgent, assert_params_close

pytestmark = pytest.mark.gpu


def make_pair(dims, B, N, path=G.PATH_AUTO, capacity=None, seed_data=synth.SEED_DATA, **kw):
    th = synth.glorot_params(dims, synth.SEED_ONLINE)
    tt = synth.glorot_params(dims, synth.SEED_TARGET)
    cap = capacity or N
    h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B,
                 replay_capacity=cap, path=path, **kw)
    h.set_params(G.NET_ONLINE, th)
    h.set_params(G.NET_TARGET, tt)
    okw = dict(loss=kw.get("loss", 0), solver=kw.get("solver", 0), delta=kw.get("huber_delta", 1.0))
    for k in ("lr", "beta2", "eps"):
        if k in kw:
            okw[k] = kw[k]
    o = OracleAgent(dims, B, th, tt, capacity=cap, **okw)
    data = synth.replay(dims, N, seed=seed_data)
    h.replay_push(*data)
    o.push(*data)
    h.debug_enable(True)
    return h, o


def check_step(h, o, idx, step_name=""):
    out, out64 = o.learn(idx)
    h.learn(idx)
    q = h.debug_fetch(G.DBG_Q_ONLINE)
    qt = h.debug_fetch(G.DBG_Q_TARGET)
    td = h.debug_fetch(G.DBG_TD)
    np.testing.assert_array_equal(h.debug_fetch(G.DBG_INDICES), idx)
    np.testing.assert_array_equal(q, out["q_online"], err_msg=f"{step_name} Q(s) not bit-exact")
    np.testing.assert_array_equal(qt, out["q_target"], err_msg=f"{step_name} Q'(s') not bit-exact")
    np.testing.assert_array_equal(td, out["td"], err_msg=f"{step_name} TD target not bit-exact")
    loss = float(h.last_loss()[0])
    assert abs(loss - out64["loss"]) <= 1e-5 * abs(out64["loss"]), (loss, out64["loss"], out["loss"])
    g = h.debug_fetch(G.DBG_GRADS)
    gs = np.abs(out64["grads"]).max()
    np.testing.assert_allclose(g, out64["grads"], rtol=1e-5, atol=1e-6 * gs, err_msg=f"{step_name} grads")
    w = h.get_params(G.NET_ONLINE)
    frac = assert_params_close(w, out64["theta"], out64["grads"], lr=o.bound_lr, what=f"{step_name} params")
    assert frac < 0.02
    m, v, it = h.get_adam_state()
    assert it == o.iter
    np.testing.assert_allclose(m, o.m, rtol=1e-4, atol=1e-6 * gs)
    # keep both sides in lock step for multi-step runs: adopt the oracle's fp32 state
    h.set_params(G.NET_ONLINE, o.theta)
    h.set_adam_state(o.m, o.v, o.iter)
    return out


@pytest.mark.parametrize("B", [20, 4096])
def test_cartpole_narrow_path(B):
    """BASELINE configs[0] (B=20) and configs[1] (B=4096): fused one-kernel step."""
    dims = synth.CARTPOLE
    N = max(4 * B, 1000)
    h, o = make_pair(dims, B, N)
    assert h.path == G.PATH_NARROW
    for step in range(3):
        idx = synth.sample_indices(N, B, step)
        check_step(h, o, idx, f"step{step}")
        if step == 1:
            h.sync_target(); o.sync_target()
            np.testing.assert_array_equal(h.get_params(G.NET_TARGET), o.theta_t)
    h.close()


@pytest.mark.parametrize("dims", [(2, 24, 24, 3), (6, 24, 24, 3), (8, 24, 24, 4)])
def test_other_classic_control_shapes(dims):
    B, N = 64, 512
    h, o = make_pair(dims, B, N)
    assert h.path == G.PATH_NARROW
    check_step(h, o, synth.sample_indices(N, B, 0))
    h.close()


@pytest.mark.parametrize("dims,B", [(synth.CARTPOLE, 20), (synth.CARTPOLE, 1000), ((5, 17, 33, 3), 37),
                                    ((128, 512, 512, 18), 256)])
def test_generic_path(dims, B):
    N = 4 * B
    h, o = make_pair(dims, B, N, path=G.PATH_GENERIC)
    assert h.path == G.PATH_GENERIC
    for step in range(2):
        idx = synth.sample_indices(N, B, step)
        check_step(h, o, idx, f"step{step}")
        # the gather itself
        s, a, _, _, _ = o.batch(idx)
        np.testing.assert_array_equal(h.debug_fetch(G.DBG_BATCH_S), s)
        np.testing.assert_array_equal(h.debug_fetch(G.DBG_BATCH_A), a)
    h.close()


@pytest.mark.parametrize("path,dims,B,delta", [(G.PATH_NARROW, synth.CARTPOLE, 512, 1.0), (G.PATH_NARROW, synth.CARTPOLE, 20, 0.25),
                                               (G.PATH_GENERIC, (5, 17, 33, 3), 37, 1.0)])
def test_pseudo_huber_loss(path, dims, B, delta):
    """model.PseudoHuber (goro loss.go:90-151, reduced by Mean) on the fp32 paths: same bars as MSE."""
    N = 4 * B
    h, o = make_pair(dims, B, N, path=path, loss=G.LOSS_PSEUDO_HUBER, huber_delta=delta)
    assert h.path == path
    for step in range(2):
        check_step(h, o, synth.sample_indices(N, B, step), f"huber step{step}")
    h.close()


@pytest.mark.parametrize("dims,B,path", [(synth.CARTPOLE, 20, G.PATH_NARROW), (synth.CARTPOLE, 20, G.PATH_GENERIC),
                                         ((5, 17, 33, 3), 37, G.PATH_GENERIC), (synth.CARTPOLE, 4096, G.PATH_NARROW)])
def test_rmsprop_solver(dims, B, path):
    """gorgonia.RMSPropSolver (solvers.go:249-335), the solver of gold's Atari config (policy.go:41-47): in the
    fused narrow kernel's tail and on the generic path (PATH_AUTO picks the fused kernel where it exists)."""
    N = 4 * B
    h, o = make_pair(dims, B, N, path=path, solver=G.SOLVER_RMSPROP, lr=1e-3, beta2=0.999, eps=1e-8)
    assert h.path == path
    for step in range(3):
        out = check_step(h, o, synth.sample_indices(N, B, step), f"rmsprop step{step}")
    m, cache, _ = h.get_adam_state()
    assert not m.any() and cache.any()
    h.close()
    h = G.Handle(batch_size=20, replay_capacity=64, solver=G.SOLVER_RMSPROP)
    assert h.path == G.PATH_NARROW             # no longer routed to the unfused path
    h.close()


def test_free_running_without_resync():
    """10 steps with NO state re-injection: drift stays within tolerance."""
    dims, B, N = synth.CARTPOLE, 256, 2048
    th = synth.glorot_params(dims, 1); tt = synth.glorot_params(dims, 2)
    h = G.Handle(batch_size=B, replay_capacity=N)
    h.set_params(G.NET_ONLINE, th); h.set_params(G.NET_TARGET, tt)
    o = OracleAgent(dims, B, th, tt)
    data = synth.replay(dims, N)
    h.replay_push(*data); o.push(*data)
    for step in range(10):
        idx = synth.sample_indices(N, B, step)
        o.learn(idx, with64=False)
        h.learn(idx)
        if step % 4 == 3:
            h.sync_target(); o.sync_target()
    w = h.get_params(G.NET_ONLINE)
    # sign-like early Adam steps amplify ulp-level gradient differences on tiny
    # gradients; bound the drift by a small multiple of lr instead of 1e-5 relative
    assert np.abs(w - o.theta).max() < 2e-4
    assert np.mean(np.abs(w - o.theta) > 1e-6) < 0.05
    q_gpu = h.predict(G.NET_ONLINE, data[0][:64])
    from oracle import binding as O
    q_cpu, _ = O.predict(dims, o.theta, data[0][:64])
    np.testing.assert_allclose(q_gpu, q_cpu, rtol=2e-3, atol=2e-4)
    h.close()


def test_replay_ring_wraparound_and_eviction():
    """Remember = PushFront + PopBack past BufferSize (agent.go:224-229)."""
    dims, B, cap = synth.CARTPOLE, 16, 100
    h, o = make_pair(dims, B, 60, path=G.PATH_GENERIC, capacity=cap)
    assert h.replay_len() == 60
    more = synth.replay(dims, 70, seed=77)
    h.replay_push(*more); o.push(*more)                  # wraps: 130 pushed into 100 slots
    assert h.replay_len() == 100 == len(o.a)
    idx = np.array([0, 99, 1, 50, 69, 70, 98, 2, 3, 4, 5, 6, 7, 8, 9, 10], np.int32)
    check_step(h, o, idx)
    s, a, _, _, _ = o.batch(idx)
    np.testing.assert_array_equal(h.debug_fetch(G.DBG_BATCH_S), s)
    # newest event is logical index 0
    np.testing.assert_array_equal(h.debug_fetch(G.DBG_BATCH_S)[0], more[0][-1])
    big = synth.replay(dims, 250, seed=78)               # one push larger than the ring
    h.replay_push(*big); o.push(*big)
    assert h.replay_len() == 100
    check_step(h, o, idx)
    h.close()


def test_device_sampler_matches_host_mirror_and_is_a_permutation():
    dims, B, N = synth.CARTPOLE, 512, 5000
    for path in (G.PATH_NARROW, G.PATH_GENERIC):
        h, o = make_pair(dims, B, N, path=path)
        h.learn(None, seed=12345)
        idx = h.debug_fetch(G.DBG_INDICES)
        np.testing.assert_array_equal(idx, G.sampler_indices(N, B, 12345))
        assert len(set(idx.tolist())) == B and idx.min() >= 0 and idx.max() < N
        # and the step it took equals the oracle's on those indices
        out, out64 = o.learn(idx)
        np.testing.assert_array_equal(h.debug_fetch(G.DBG_TD), out["td"])
        h.close()
    full = G.sampler_indices(1000, 1000, 7)
    assert sorted(full.tolist()) == list(range(1000))    # a bijection on [0, len)


def test_predict_and_greedy_action_bit_exact():
    from oracle import binding as O
    for dims in (synth.CARTPOLE, (128, 512, 512, 18)):
        th = synth.glorot_params(dims, 3)
        h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3],
                     batch_size=8, replay_capacity=64, path=G.PATH_GENERIC)
        h.set_params(G.NET_ONLINE, th)
        x = synth.replay(dims, 300, seed=5)[0]
        q_ref, g_ref = O.predict(dims, th, x)
        np.testing.assert_array_equal(h.predict(G.NET_ONLINE, x), q_ref)
        np.testing.assert_array_equal(h.greedy_action(x), g_ref)
        np.testing.assert_array_equal(h.predict(G.NET_ONLINE, x[:1]), q_ref[:1])   # Predict: one row
        h.close()


def test_greedy_action_tie_and_nan_rules():
    dims = (1, 1, 1, 4)
    # q = b3 regardless of x when W3 = 0
    def run(b3):
        th = np.array([1, 0, 1, 0, 0, 0, 0, 0] + list(b3), np.float32)
        h = G.Handle(obs_dim=1, hidden1=1, hidden2=1, n_actions=4, batch_size=1, replay_capacity=4,
                     path=G.PATH_GENERIC)
        h.set_params(G.NET_ONLINE, th)
        a = int(h.greedy_action(np.array([[1.0]], np.float32))[0])
        h.close()
        return a
    assert run([1, 3, 3, 2]) == 1
    assert run([9, np.nan, 10, 0]) == 1
    assert run([9, 1, np.inf, 10]) == 2
    assert run([np.nan, 1, 2, 0]) == 0


@pytest.mark.parametrize("dims,B", [(synth.CARTPOLE, 20), ((16, 48, 32, 6), 96)])
def test_fit_batch_matches_oracle(dims, B):
    """goro Sequential.FitBatch / Fit on explicit labels (model.go:528-573)."""
    from oracle import binding as O
    th = synth.glorot_params(dims, 1)
    P = synth.nparams(dims)
    h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B,
                 replay_capacity=64)
    h.set_params(G.NET_ONLINE, th)
    rng = np.random.Generator(np.random.PCG64(9))
    w = th.astype(np.float64); m = np.zeros(P); v = np.zeros(P); it = 0
    for n in (B, 1, B):
        x = rng.uniform(-1, 1, (n, dims[0])).astype(np.float32)
        y = rng.uniform(-1, 1, (n, dims[3])).astype(np.float32)
        h.fit_batch(x, y)
        ref = O.fit(dims, w, m, v, it, x.astype(np.float64), y.astype(np.float64), dtype=np.float64)
        it = ref["iter"]
        loss = float(h.last_loss()[0])
        assert abs(loss - ref["loss"]) <= 1e-5 * abs(ref["loss"])
        g = h.debug_fetch(G.DBG_GRADS)
        np.testing.assert_allclose(g, ref["grads"], rtol=1e-5, atol=1e-6 * np.abs(ref["grads"]).max())
        assert_params_close(h.get_params(G.NET_ONLINE), w, ref["grads"])
        h.set_params(G.NET_ONLINE, w.astype(np.float32))
        w = w.astype(np.float32).astype(np.float64)
        mm, vv, _ = h.get_adam_state()
        m = mm.astype(np.float64); v = vv.astype(np.float64)
    with pytest.raises(G.GoldqError):
        h.fit_batch(np.zeros((B + 1, dims[0]), np.float32), np.zeros((B + 1, dims[3]), np.float32))
    h.close()


def test_replicas_are_independent_agents():
    """BASELINE configs[4]: independent CartPole agents batched in one grid."""
    dims, B, N, R = synth.CARTPOLE, 20, 256, 5
    P = synth.nparams(dims)
    ths = [synth.glorot_params(dims, 10 + r) for r in range(R)]
    tts = [synth.glorot_params(dims, 20 + r) for r in range(R)]
    h = G.Handle(batch_size=B, replay_capacity=N, n_replicas=R)
    assert h.path == G.PATH_NARROW
    h.set_params(G.NET_ONLINE, np.concatenate(ths)); h.set_params(G.NET_TARGET, np.concatenate(tts))
    h.debug_enable(True)
    os_ = [OracleAgent(dims, B, ths[r], tts[r]) for r in range(R)]
    datas = [synth.replay(dims, N, seed=100 + r) for r in range(R)]
    h.replay_push(*[np.concatenate([d[k] for d in datas]) for k in range(5)])
    for r in range(R):
        os_[r].push(*datas[r])
    for step in range(2):
        idx = np.stack([synth.sample_indices(N, B, 50 * r + step) for r in range(R)])
        h.learn(idx)
        q = h.debug_fetch(G.DBG_Q_ONLINE).reshape(R, B, -1)
        td = h.debug_fetch(G.DBG_TD).reshape(R, B)
        g = h.debug_fetch(G.DBG_GRADS).reshape(R, P)
        w = h.get_params(G.NET_ONLINE).reshape(R, P)
        loss = h.last_loss()
        for r in range(R):
            out, out64 = os_[r].learn(idx[r])
            np.testing.assert_array_equal(q[r], out["q_online"])
            np.testing.assert_array_equal(td[r], out["td"])
            np.testing.assert_allclose(g[r], out64["grads"], rtol=1e-5, atol=1e-6 * np.abs(out64["grads"]).max())
            assert abs(loss[r] - out64["loss"]) <= 1e-5 * abs(out64["loss"])
            assert_params_close(w[r], out64["theta"], out64["grads"])
        h.set_params(G.NET_ONLINE, np.concatenate([o.theta for o in os_]))
        h.set_adam_state(np.concatenate([o.m for o in os_]), np.concatenate([o.v for o in os_]), os_[0].iter)
    # greedy actions per replica
    x = datas[0][0][:32]
    from oracle import binding as O
    for r in (0, R - 1):
        np.testing.assert_array_equal(h.greedy_action(x, replica=r), O.predict(dims, os_[r].theta, x)[1])
    h.close()


def test_error_behaviour():
    h = G.Handle(batch_size=20, replay_capacity=100)
    with pytest.raises(G.GoldqError) as e:           # Len < batch: agent.go:127 (host layer maps to a no-op)
        h.learn(np.zeros(20, np.int32))
    assert e.value.code == G.ENOTREADY
    h.replay_push(*synth.replay(synth.CARTPOLE, 30))
    with pytest.raises(G.GoldqError) as e:
        h.learn(np.full(20, 30, np.int32))           # index == Len is out of range (deque.At panics)
    assert e.value.code == G.EINVAL
    with pytest.raises(G.GoldqError):
        G.Handle(batch_size=0)
    with pytest.raises(G.GoldqError):
        G.Handle(obs_dim=5, path=G.PATH_NARROW)      # shape not instantiated: fail loudly, no fallback
    with pytest.raises(G.GoldqError):
        G.Handle(n_replicas=2, obs_dim=5)            # replicas need the narrow path
    h.close()


def test_large_batch_linearity_property():
    """Size-independent property at full size: with all events terminal (target = r)
    the gradient is linear in the TD error, so g(B rows) == mean of g over two halves."""
    dims, B = synth.CARTPOLE, 4096
    th = synth.glorot_params(dims, 1)
    s, a, r, s2, _ = synth.replay(dims, B)
    done = np.ones(B, np.uint8)
    def grads(rows):
        n = len(rows)
        h = G.Handle(batch_size=n, replay_capacity=n)
        h.set_params(G.NET_ONLINE, th); h.set_params(G.NET_TARGET, th)
        h.replay_push(s[rows], a[rows], r[rows], s2[rows], done[rows])
        h.learn(np.arange(n, dtype=np.int32))
        g = h.debug_fetch(G.DBG_GRADS).astype(np.float64); h.close()
        return g
    g_all = grads(np.arange(B)); g_a = grads(np.arange(B // 2)); g_b = grads(np.arange(B // 2, B))
    np.testing.assert_allclose(g_all, 0.5 * (g_a + g_b), rtol=1e-4, atol=1e-6 * np.abs(g_all).max())


@pytest.mark.parametrize("kind", ["narrow", "generic", "wide", "replicas"])
def test_learn_many_graph_replay_equals_single_steps(kind):
    """goldq_learn_many (CUDA-graph replay, per-step scalars in device memory) must be
    bit-identical to the same steps issued one by one, across graph rebuilds."""
    if kind == "wide":
        dims, B, N, kw = (64, 256, 128, 6), 256, 2048, dict(precision=G.PREC_BF16)
    elif kind == "generic":
        dims, B, N, kw = (5, 17, 33, 3), 64, 512, dict(path=G.PATH_GENERIC)
    elif kind == "replicas":
        dims, B, N, kw = synth.CARTPOLE, 20, 256, dict(n_replicas=3)
    else:
        dims, B, N, kw = synth.CARTPOLE, 512, 4096, {}
    R = kw.get("n_replicas", 1)
    th = np.concatenate([synth.glorot_params(dims, 1 + r) for r in range(R)])
    tt = np.concatenate([synth.glorot_params(dims, 40 + r) for r in range(R)])
    parts = [synth.replay(dims, N, seed=5 + r) for r in range(R)]
    data = [np.concatenate([p[k] for p in parts]) for k in range(5)]
    more = [synth.replay(dims, 100, seed=70 + r) for r in range(R)]
    more = [np.concatenate([p[k] for p in more]) for k in range(5)]

    def run(mode):
        h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B,
                     replay_capacity=N, **kw)
        h.set_params(G.NET_ONLINE, th); h.set_params(G.NET_TARGET, tt)
        h.replay_push(*data)

        def steps(n, seed):
            if mode == "many":                     # graphs of 32/28/../4/2/1 steps
                h.learn_many(n, seed)
            elif mode == "single":                 # the one-step graph, n times
                for i in range(n):
                    h.learn(None, seed + i)
            else:                                  # eager launches (the profiling entry runs a real step)
                for i in range(n):
                    h.step_breakdown(seed + i)

        seed = 1000
        for n in (5, 37, 2):                       # 37 = 32 + 4 + 1 (graph sizes 32, 28, .. 4, 2, 1)
            steps(n, seed)
            seed += n
            h.sync_target()
        builds = h.graph_builds()
        h.replay_push(*more)                       # moves the ring: the step slots carry head/len, no rebuild
        steps(4, seed)
        assert h.graph_builds() == builds, "a replay push invalidated the step graphs"
        w = h.get_params(G.NET_ONLINE); m, v, it = h.get_adam_state(); loss = h.last_loss()
        h.close()
        return w, m, v, it, loss

    a, b, c = run("single"), run("many"), run("eager")
    assert a[3] == b[3] == c[3] == 48
    for other in (b, c):
        for x, y in zip(a[:3], other[:3]):
            np.testing.assert_array_equal(x, y)
        np.testing.assert_array_equal(a[4], other[4])


@pytest.mark.parametrize("kind", ["narrow", "wide"])
def test_remember_learn_loop_never_rebuilds_graphs(kind):
    """Agent.Learn's real loop -- Remember one event, Learn once -- and learn_many between pushes replay
    the graphs instantiated by goldq_learn_prepare: zero instantiations afterwards, and the ring
    position that travels in the step slot selects the rows the host mirror of the sampler predicts."""
    if kind == "wide":
        dims, B, N, kw = (64, 256, 128, 6), 256, 1024, dict(precision=G.PREC_BF16)
    else:
        dims, B, N, kw = synth.CARTPOLE, 64, 300, {}
    h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B,
                 replay_capacity=N, **kw)
    h.set_params(G.NET_ONLINE, synth.glorot_params(dims, 1)); h.set_params(G.NET_TARGET, synth.glorot_params(dims, 2))
    h.replay_push(*synth.replay(dims, N - 50, seed=9))
    h.debug_enable(True)
    h.learn_prepare()
    builds = h.graph_builds()
    assert builds == 11                             # 32, 28, .. 8, 4, 2, 1 steps + the host-index step
    ev = synth.replay(dims, 120, seed=10)
    for i in range(120):                            # wraps the ring (capacity N)
        h.replay_push(*[x[i:i + 1] for x in ev])
        if i % 3 == 0:
            h.learn(None, 5000 + i)
            want = G.sampler_indices(h.replay_len(), B, 5000 + i)
            np.testing.assert_array_equal(h.debug_fetch(G.DBG_INDICES), want)
        elif i % 3 == 1:
            h.learn_many(7, 9000 + 7 * i)
        else:
            h.learn(synth.sample_indices(h.replay_len(), B, i))
    h.sync()
    assert h.graph_builds() == builds
    assert np.isfinite(h.last_loss()).all()
    h.close()


def test_packed_push_equals_unpacked_and_bad_actions_are_rejected():
    dims, B, N = synth.CARTPOLE, 32, 256
    data = synth.replay(dims, 200, seed=3)
    hs = []
    for packed in (False, True):
        h = G.Handle(batch_size=B, replay_capacity=N)
        h.set_params(G.NET_ONLINE, synth.glorot_params(dims, 1)); h.set_params(G.NET_TARGET, synth.glorot_params(dims, 2))
        for lo in (0, 77, 150):                     # ragged chunk sizes
            hi = min(200, lo + 77) if lo else 77
            chunk = [x[lo:hi] for x in data]
            if packed:
                buf = h.pack(*chunk)
                assert buf.size == h.packed_bytes(hi - lo)
                h.replay_push_packed(hi - lo, buf.ctypes.data)
                h.sync()
            else:
                h.replay_push(*chunk)
        h.debug_enable(True)
        h.learn(synth.sample_indices(h.replay_len(), B, 1))
        hs.append((h.debug_fetch(G.DBG_Q_ONLINE), h.debug_fetch(G.DBG_TD), h.get_params(G.NET_ONLINE), h))
    for x, y in zip(hs[0][:3], hs[1][:3]):
        np.testing.assert_array_equal(x, y)
    h = hs[0][3]
    n0 = h.replay_len()
    bad = [x[:5].copy() for x in data]
    for wrong in (-1, dims[3]):                     # the reference panics in qValues.Set (agent.go:155)
        bad[1][3] = wrong
        with pytest.raises(G.GoldqError) as e:
            h.replay_push(*bad)
        assert e.value.code == G.EINVAL and h.replay_len() == n0
    for _, _, _, hh in hs:
        hh.close()


def test_staged_push_commits_in_order_and_leaves_the_memory_alone_until_then():
    """goldq_replay_stage_packed starts the copy only; goldq_replay_commit appends the oldest staged batch; a third
    stage is refused; a later push commits what is still staged first.  The resulting memory (and a learn step and
    its loss on it) equals three plain pushes; a learn step enqueued BEFORE the commit does not see the batch."""
    dims, B, N = synth.CARTPOLE, 32, 256
    data = synth.replay(dims, 180, seed=9)
    chunks = [[x[lo:lo + 60] for x in data] for lo in (0, 60, 120)]
    idx = synth.sample_indices(60, B, 1)
    out = []
    for staged in (False, True):
        h = G.Handle(batch_size=B, replay_capacity=N)
        h.set_params(G.NET_ONLINE, synth.glorot_params(dims, 1)); h.set_params(G.NET_TARGET, synth.glorot_params(dims, 2))
        bufs = [h.pack(*c) for c in chunks]
        if staged:
            h.replay_push_packed(60, bufs[0].ctypes.data)
            h.replay_stage_packed(60, bufs[1].ctypes.data)
            assert h.replay_len() == 60
            h.learn(idx)                              # before the commit: trains on the first batch only
            h.replay_stage_packed(60, bufs[2].ctypes.data)
            with pytest.raises(G.GoldqError) as e:
                h.replay_stage_packed(60, bufs[0].ctypes.data)
            assert e.value.code == G.EINVAL
            h.replay_commit()
            assert h.replay_len() == 120
            h.replay_commit()
            assert h.replay_len() == 180
            with pytest.raises(G.GoldqError) as e:
                h.replay_commit()
            assert e.value.code == G.EINVAL
        else:
            h.replay_push_packed(60, bufs[0].ctypes.data)
            h.learn(idx)
            h.replay_push_packed(60, bufs[1].ctypes.data)
            h.replay_push_packed(60, bufs[2].ctypes.data)
        l1 = h.last_loss().copy()
        h.learn(synth.sample_indices(180, B, 2))
        out.append((l1, h.last_loss().copy(), h.get_params(G.NET_ONLINE)))
        # a push with something still staged: order is kept
        h.replay_stage_packed(60, bufs[0].ctypes.data)
        h.replay_push_packed(60, bufs[1].ctypes.data)
        assert h.replay_len() == 256
        h.sync()
        h.close()
    for x, y in zip(out[0], out[1]):
        np.testing.assert_array_equal(x, y)
    assert np.isfinite(out[0][0]).all() and out[0][0][0] != out[0][1][0]


def test_batched_epsilon_greedy_action():
    """goldq_epsilon_greedy_action: rows that do not explore take the oracle's first-argmax, rows that
    explore take the planned uniform action; epsilon 0 is goldq_greedy_action."""
    from oracle import binding as O
    dims, n = (5, 17, 33, 3), 4000
    th = synth.glorot_params(dims, 3)
    h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=8, replay_capacity=64)
    h.set_params(G.NET_ONLINE, th)
    x = np.random.default_rng(5).uniform(-1, 1, (n, dims[0])).astype(np.float32)
    _, greedy = O.predict(dims, th, x)
    np.testing.assert_array_equal(h.greedy_action(x), greedy)
    np.testing.assert_array_equal(h.epsilon_greedy_action(x, 0.0, 99), greedy)
    act = h.epsilon_greedy_action(x, 0.3, 4242)
    ex, ra = G.explore_plan(n, dims[3], 0.3, 4242)
    assert 0.25 < ex.mean() < 0.35
    np.testing.assert_array_equal(act[~ex], greedy[~ex])
    np.testing.assert_array_equal(act[ex], ra[ex])
    h.close()
