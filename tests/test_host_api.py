"""The host layer mirrors gold's deepq package; these tests read like the reference's own
(pkg/v1/agent/deepq/memory_test.go, policy_test.go) with assertions added where the
reference only logs."""
import numpy as np
import pytest

from gold_b200 import deepq
from gold_b200 import synth


def test_memory():
    # memory_test.go:12-29: push 10 events, Sample(5) returns 5
    mem = deepq.Memory()
    for i in range(10):
        s = np.array([[i, i + 1, i + 2, i + 3]], np.float32)
        ev = deepq.new_event(s, 0, deepq.Outcome(observation=s, action=i, reward=1.0, done=False))
        mem.push_front(ev)
    sample = mem.sample(5)
    assert len(sample) == 5
    assert len({e.i for e in sample}) == 5                       # distinct, like rand.Perm(Len)[:5]
    for e in sample:
        assert mem.at(e.i) is e                                   # event.i is its deque index (memory.go:62-63)
    assert mem.at(0).outcome.action == 9                          # PushFront: newest at index 0
    with pytest.raises(ValueError, match="queue size 10 is less than batch size 11"):
        mem.sample(11)                                            # memory.go:55-57


def test_decay_schedule():
    # schedule.go:101-105,130-141
    s = deepq.default_decay_schedule()
    assert s.initial() == 1.0
    assert s.value() == float(np.float32(0.995))
    v = [s.value() for _ in range(2000)]
    assert v[-1] == float(np.float32(0.01)) and min(v) == v[-1]
    s2 = deepq.default_decay_schedule(decay_rate=0.9995)          # experiments/cartpole/main.go:28
    assert abs(s2.value() - 0.9995) < 1e-7


def test_unsupported_graphs_fail_loudly():
    env = deepq.ShapesEnv(4, 2)
    bad = deepq.PolicyConfig(layer_builder=lambda x, y: [deepq.FC(x, 24), deepq.FC(24, y, activation=deepq.LINEAR)])
    with pytest.raises(ValueError, match="no CPU fallback"):
        deepq.Agent(deepq.AgentConfig(policy_config=bad), env)
    with pytest.raises(ValueError, match="environment cannot be nil"):
        deepq.Agent(deepq.AgentConfig(), None)


@pytest.mark.gpu
def test_policy_converges_to_static_values():
    """policy_test.go:16-83 ("test that network converges to static values"), with the
    final predictions asserted instead of logged."""
    env = deepq.ShapesEnv(4, 2)
    agent = deepq.Agent(deepq.AgentConfig(seed=7), env)
    m = agent.policy
    x1 = np.array([[0.051960364, 0.14512223, 0.12799974, -2.0140305]], np.float32)
    x2 = np.array([[0.15163246, -0.94560495, 0.3904586, -8.20394809]], np.float32)
    y1 = np.array([[0.4484117, 0.09160687]], np.float32)
    y2 = np.array([[3.23904234, 2.203948023]], np.float32)
    q0 = m.predict(x1)
    assert q0.shape == (1, 2)
    for i in range(3000):
        m.fit(x1, y1)
        m.fit(x2, y2)
    np.testing.assert_allclose(m.predict(x1), y1, atol=0.05)
    np.testing.assert_allclose(m.predict(x2), y2, atol=0.05)
    with pytest.raises(ValueError):
        agent.target_policy.fit(x1, y1)
    agent.close()


@pytest.mark.gpu
def test_agent_loop_semantics():
    """Action / Remember / Learn cadence of experiments/cartpole/main.go:27-58 on synthetic
    transitions: Learn is a no-op until Len >= batch (agent.go:127), epsilon decays once per
    Learn (agent.go:171), steps counts Actions only and the target is cloned when
    steps % UpdateTargetSteps == 0 (agent.go:181-190)."""
    env = deepq.ShapesEnv(4, 2, seed=1)
    hp = deepq.Hyperparameters(update_target_steps=7, buffer_size=500)
    agent = deepq.Agent(deepq.AgentConfig(hyperparameters=hp, seed=3), env)
    s, a, r, s2, d = synth.replay(synth.CARTPOLE, 120, seed=11)
    target0 = agent.target_policy.learnables().copy()
    eps = [agent.epsilon]
    synced_at = []
    for t in range(120):
        act = agent.action(s[t:t + 1])
        assert act in (0, 1)
        agent.remember(deepq.new_event(s[t:t + 1], act, deepq.Outcome(s2[t:t + 1], act, r[t], bool(d[t]))))
        before = agent.target_policy.learnables().copy()
        assert agent.learn() is None
        after = agent.target_policy.learnables()
        if not np.array_equal(before, after):
            synced_at.append(agent.steps)
        eps.append(agent.epsilon)
    assert agent.steps == 120
    # no learning (and no epsilon decay) before 20 events are in memory
    assert eps[19] == eps[0] == 1.0 and eps[20] < 1.0
    assert abs(eps[-1] - 0.995 ** 101) < 1e-5
    assert synced_at and all(st % 7 == 0 for st in synced_at) and synced_at[0] == 21
    np.testing.assert_array_equal(agent.target_policy.learnables(), agent.policy.learnables()) if agent.steps % 7 == 0 else None
    assert not np.array_equal(target0, agent.target_policy.learnables())
    assert np.isfinite(agent.last_loss())
    agent.close()


@pytest.mark.gpu
def test_her_learn_step_is_deepq_on_state_goal():
    """her/agent.go:129-189: the learn step equals the DQN step on state‖goal inputs —
    bit-exact Q/TD against the oracle fed with the concatenated tensors."""
    from gold_b200 import _capi as G, her
    from oracle import binding as O
    sd, gd, A, B = 2, 2, 2, 20          # goals live in observation space (her/agent.go:251); obs_dim = 4
    agent = her.Agent(her.AgentConfig(seed=5), sd, gd, A)
    rng = np.random.Generator(np.random.PCG64(1))
    events = [her.Event(rng.uniform(-1, 1, sd), rng.uniform(-1, 1, gd), int(rng.integers(0, A)),
                        rng.uniform(-1, 1, sd), -1.0, False) for _ in range(64)]
    agent.remember(*events)
    th = agent.policy.learnables().copy(); tt = agent.target_policy.learnables().copy()
    agent._h.debug_enable(True)
    idx = np.arange(20, dtype=np.int32) * 3
    agent.learn(indices=idx)
    # her.Memory indexes its events oldest first (events[value], her/memory.go:75-117)
    rows = [events[i] for i in idx]
    s = np.stack([np.concatenate([e.state, e.goal]) for e in rows])
    s2 = np.stack([np.concatenate([e.observation, e.goal]) for e in rows])
    P = th.size
    w = th.copy()
    ref = O.learn((4, 24, 24, 2), B, w, tt, np.zeros(P, np.float32), np.zeros(P, np.float32), 0, s,
                  [e.action for e in rows], [e.reward for e in rows], s2, [e.done for e in rows], gamma=0.9)
    np.testing.assert_array_equal(agent._h.debug_fetch(G.DBG_Q_ONLINE), ref["q_online"])
    np.testing.assert_array_equal(agent._h.debug_fetch(G.DBG_TD), ref["td"])
    np.testing.assert_allclose(agent.policy.learnables(), w, rtol=1e-5, atol=2e-6)
    assert agent.episodes == 1
    # hindsight: one Learn per relabelled event, last one terminal with the successful reward
    n0 = agent.memory_len()
    agent.hindsight(events[:5])
    assert agent.memory_len() == n0 + 5 and agent.episodes == 6
    a = agent.action(events[0].state, events[0].goal)
    assert a in (0, 1) and agent.epsilon < 1.0
    agent.close()


@pytest.mark.gpu
def test_her_accepts_pseudo_huber_and_rmsprop_like_deepq():
    """her.Agent shares deepq's loss / solver mapping: a PseudoHuber + RMSProp config reaches the device as
    such (it used to train with MSE silently / fail on .beta1), and unknown kinds fail loudly."""
    from gold_b200 import _capi as G, her
    pc = deepq.PolicyConfig(loss=deepq.PseudoHuberLoss(0.5), optimizer=deepq.RMSPropSolver(learn_rate=1e-3), batch_size=8)
    agent = her.Agent(her.AgentConfig(policy_config=pc, seed=1), 2, 2, 2)
    assert agent._h.cfg.loss == G.LOSS_PSEUDO_HUBER and agent._h.cfg.solver == G.SOLVER_RMSPROP
    assert abs(agent._h.cfg.huber_delta - 0.5) < 1e-7
    agent.close()
    with pytest.raises(ValueError):
        her.Agent(her.AgentConfig(policy_config=deepq.PolicyConfig(optimizer=object())), 2, 2, 2)


@pytest.mark.gpu
def test_batched_acting_syncs_the_target_when_a_multiple_is_stepped_over():
    """Agent.updateTarget (agent.go:181-190) with action_batch: 64 environments and a period of 100 used to sync
    only every 1600 steps; a multiple stepped over between two Learns now triggers the sync."""
    hp = deepq.Hyperparameters(update_target_steps=100)
    agent = deepq.Agent(deepq.AgentConfig(hyperparameters=hp, policy_config=deepq.PolicyConfig(batch_size=8), seed=2), deepq.ShapesEnv(4, 2))
    rng = np.random.Generator(np.random.PCG64(0))
    for _ in range(16):
        agent.remember(deepq.new_event(rng.uniform(-1, 1, 4).astype(np.float32), 0,
                                   deepq.Outcome(rng.uniform(-1, 1, 4).astype(np.float32), 0, 1.0, False)))
    states = rng.uniform(-1, 1, (64, 4)).astype(np.float32)
    synced = []
    for it in range(4):                           # steps: 64, 128, 192, 256 -> multiples 100 and 200 are stepped over
        agent.action_batch(states)
        before = agent.target_policy.learnables().copy()
        agent.learn()
        synced.append(not np.array_equal(before, agent.target_policy.learnables()))
    assert synced == [False, True, False, True]
    agent.close()


@pytest.mark.gpu
def test_resize_batch_keeps_parameters_and_learns_at_the_new_size():
    """Sequential.ResizeBatch (goro model.go:485-497) on the narrow and the wide path: same parameters, solver
    state and replay; the next step runs at the new batch size and matches the oracle."""
    from gold_b200 import _capi as G, synth
    from util import OracleAgent
    for dims, B0, B1, kw, tol in ((synth.CARTPOLE, 20, 64, {}, 1e-5), ((64, 256, 128, 6), 256, 512, dict(precision=G.PREC_BF16), 2e-2)):
        N = 1024
        th, tt = synth.glorot_params(dims, 1), synth.glorot_params(dims, 2)
        h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B0, replay_capacity=N, **kw)
        h.set_params(G.NET_ONLINE, th); h.set_params(G.NET_TARGET, tt)
        data = synth.replay(dims, N)
        h.replay_push(*data)
        o = OracleAgent(dims, B0, th, tt)
        o.push(*data)
        idx = synth.sample_indices(N, B0, 0)
        h.learn(idx); o.learn(idx, with64=False)
        w_before = h.get_params(G.NET_ONLINE)
        h.resize_batch(B1)
        assert h.B == B1 and h.replay_len() == N
        np.testing.assert_array_equal(h.get_params(G.NET_ONLINE), w_before)
        o.B = B1
        o.theta[:] = w_before                      # (the bf16 path's parameters differ from the fp32 oracle's by its tolerance)
        m, v, it = h.get_adam_state(); o.m[:] = m; o.v[:] = v; o.iter = it
        idx = synth.sample_indices(N, B1, 1)
        out, _ = o.learn(idx, with64=False)
        h.learn(idx)
        loss = float(h.last_loss()[0])
        assert abs(loss - out["loss"]) <= tol * abs(out["loss"]), (dims, loss, out["loss"])
        with pytest.raises((G.GoldqError, AssertionError)):
            h.learn(synth.sample_indices(N, B0, 2))        # the old batch size is now a shape error for the index list
        h.close()


@pytest.mark.gpu
def test_atari_policy_config_builds_the_conv_agent():
    """deepq.DefaultAtariPolicyConfig (policy.go:41-47, 62-71) through the host layer: the conv stack is recognised,
    runs on GOLDQ_PATH_CONV with RMSProp, learns, acts; a different convolutional graph fails loudly."""
    from gold_b200 import _capi as G
    env = deepq.ShapesEnv((1, 84, 84), 6, seed=3)
    agent = deepq.Agent(deepq.AgentConfig(hyperparameters=deepq.Hyperparameters(buffer_size=64),
                                          policy_config=deepq.default_atari_policy_config(), seed=4), env)
    assert agent._h.path == G.PATH_CONV and agent._h.cfg.solver == G.SOLVER_RMSPROP and agent.batch_size == 20
    rng = np.random.Generator(np.random.PCG64(0))
    for t in range(24):
        s = rng.uniform(0, 1, (1, 84, 84)).astype(np.float32)
        a = agent.action(s)
        assert 0 <= a < 6
        agent.remember(deepq.new_event(s, a, deepq.Outcome(rng.uniform(0, 1, (1, 84, 84)).astype(np.float32), a, 1.0, t % 7 == 6)))
        agent.learn()
    assert np.isfinite(agent.last_loss()) and agent.last_loss() > 0
    q = agent.policy.predict(s)
    assert q.shape == (1, 6) and np.isfinite(q).all()
    agent.close()
    bad = deepq.PolicyConfig(layer_builder=lambda x, y: [deepq.Conv2D(1, 16, 8, 8, stride=(4, 4)), deepq.Flatten(), deepq.FC(64, y, activation=deepq.LINEAR)])
    with pytest.raises(ValueError):
        deepq.Agent(deepq.AgentConfig(policy_config=bad), env)
