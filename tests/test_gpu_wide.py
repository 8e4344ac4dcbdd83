"""bf16 tcgen05 path (BASELINE configs[2]) against the fp32/fp64 oracle.

Stated tolerance (north_star: "bf16 wide configs within a stated 1e-2"):
  Q(s), Q'(s'), TD targets : |x - ref| <= 1e-2 * max|ref|      (bf16 operands, fp32 accumulate)
  loss                      : 2e-2 relative
  gradients                 : ||g - ref||_2 <= 2e-2 ||ref||_2 and elementwise 2e-2 * max|ref|
  post-Adam parameters      : 1e-5 relative on well-conditioned elements (|g| >= 1e-2 max|g|),
                              hard bound 2*lr*(1+|g|) on the rest (tests/util.py)
  greedy actions            : equal wherever the oracle's top-2 Q gap exceeds the Q tolerance
"""
import numpy as np
import pytest

from gold_b200 import _capi as G
from gold_b200 import synth
from util import OracleAgent, assert_params_close

pytestmark = pytest.mark.gpu


def make(dims, B, N):
    th = synth.glorot_params(dims, 1); tt = synth.glorot_params(dims, 2)
    h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B,
                 replay_capacity=N, precision=G.PREC_BF16)
    assert h.path == G.PATH_WIDE
    h.set_params(G.NET_ONLINE, th); h.set_params(G.NET_TARGET, tt)
    o = OracleAgent(dims, B, th, tt)
    data = synth.replay(dims, N)
    h.replay_push(*data); o.push(*data)
    h.debug_enable(True)
    return h, o


def check(h, o, idx):
    out, out64 = o.learn(idx)
    h.learn(idx)
    q = h.debug_fetch(G.DBG_Q_ONLINE); qt = h.debug_fetch(G.DBG_Q_TARGET); td = h.debug_fetch(G.DBG_TD)
    qs = np.abs(out64["q_online"]).max()
    assert np.abs(q - out64["q_online"]).max() <= 1e-2 * qs, np.abs(q - out64["q_online"]).max() / qs
    nd = out["done"] == 0
    assert np.abs(qt[nd] - out64["q_target"][nd]).max() <= 1e-2 * np.abs(out64["q_target"]).max()
    assert (qt[~nd] == 0).all()
    assert np.abs(td - out64["td"]).max() <= 1e-2 * np.abs(out64["td"]).max()
    loss = float(h.last_loss()[0])
    assert abs(loss - out64["loss"]) <= 2e-2 * abs(out64["loss"]), (loss, out64["loss"])
    g = h.debug_fetch(G.DBG_GRADS).astype(np.float64)
    ref = out64["grads"]
    assert np.linalg.norm(g - ref) <= 2e-2 * np.linalg.norm(ref), np.linalg.norm(g - ref) / np.linalg.norm(ref)
    assert np.abs(g - ref).max() <= 2e-2 * np.abs(ref).max()
    # per-block checks so that a broken small block (biases, W3) cannot hide in the norm
    off = synth.offsets(o.dims)
    for name, (a, b) in off.items():
        nr = np.linalg.norm(ref[a:b])
        assert np.linalg.norm(g[a:b] - ref[a:b]) <= 3e-2 * nr + 1e-12, (name, np.linalg.norm(g[a:b] - ref[a:b]) / nr)
    w = h.get_params(G.NET_ONLINE)
    frac = assert_params_close(w, out64["theta"], ref, lr=o.lr, cond=1e-2, what="bf16 params")
    assert frac < 0.25
    h.set_params(G.NET_ONLINE, o.theta)
    h.set_adam_state(o.m, o.v, o.iter)


@pytest.mark.parametrize("dims,B", [(synth.WIDE, 256), (synth.WIDE, 8192), ((64, 256, 128, 6), 200)])
def test_wide_bf16_learn_step(dims, B):
    N = max(2 * B, 1024)
    h, o = make(dims, B, N)
    for step in range(2):
        check(h, o, synth.sample_indices(N, B, step))
        if step == 0:
            h.sync_target(); o.sync_target()
    h.close()


def test_wide_greedy_actions_match_outside_ties():
    from oracle import binding as O
    dims = synth.WIDE
    h, o = make(dims, 256, 1024)
    x = synth.replay(dims, 512, seed=3)[0]
    q64 = O.predict64(dims, o.theta.astype(np.float64), x.astype(np.float64))
    a_gpu = h.greedy_action(x)            # predict runs the fp32 reference-arithmetic kernels
    q32, a_ref = O.predict(dims, o.theta, x)
    np.testing.assert_array_equal(a_gpu, a_ref)
    h.close()


def test_wide_free_running_loss_tracks_oracle():
    dims, B, N = synth.WIDE, 1024, 4096
    h, o = make(dims, B, N)
    lg, lo = [], []
    for step in range(6):
        idx = synth.sample_indices(N, B, step)
        out, _ = o.learn(idx, with64=False)
        h.learn(idx)
        lg.append(float(h.last_loss()[0])); lo.append(float(out["loss"]))
        if step == 2:
            h.sync_target(); o.sync_target()
    np.testing.assert_allclose(lg, lo, rtol=3e-2)
    h.close()
