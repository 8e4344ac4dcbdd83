"""gold's Atari conv policy (policy.go:41-47,62-71) on the device (GOLDQ_PATH_CONV, fp32) against the oracle
(oracle/conv_core.inc, itself checked against torch autograd in tests/test_oracle_conv.py).
Bars: Q(s), Q'(s'), TD targets bit-exact (the forward reproduces gorgonia's im2col . filter^T with gonum's
DotUnitary lane order, the FC layers its k-sequential Sgemm); loss and gradients 1e-5 relative against the fp64
twin; parameters after the solver step by the shared rule (tests/util.py)."""
import numpy as np
import pytest

from gold_b200 import _capi as G
from gold_b200 import synth
from oracle import binding as O
from util import assert_params_close

pytestmark = pytest.mark.gpu


def _params(P, seed, C, H, W, A):
    """Glorot-like scales per block so that activations stay O(1) through the stack."""
    rng = np.random.Generator(np.random.PCG64(seed))
    F = O.atari_features(C, H, W)
    blocks = [(32 * C * 64, C * 64), (64 * 32 * 16, 32 * 16), (64 * 64 * 9, 64 * 9), (F * 512, F), (512, 1), (512 * A, 512), (A, 1)]
    out = np.concatenate([rng.standard_normal(n) * np.sqrt(2.0 / (fan + 32)) for n, fan in blocks]).astype(np.float32)
    assert out.size == P
    return out


def _make(shape, A, B, N, solver=G.SOLVER_ADAM, **kw):
    C, H, W = shape
    D = C * H * W
    h = G.Handle(arch=G.ARCH_ATARI_CONV, in_channels=C, in_height=H, in_width=W, obs_dim=D, n_actions=A, batch_size=B,
                 replay_capacity=N, solver=solver, **kw)
    assert h.path == G.PATH_CONV and h.P == O.atari_nparams(C, H, W, A)
    th, tt = _params(h.P, 1, C, H, W, A), _params(h.P, 2, C, H, W, A)
    h.set_params(G.NET_ONLINE, th); h.set_params(G.NET_TARGET, tt)
    data = synth.replay((D, 0, 0, A), N, seed=11)
    h.replay_push(*data)
    h.debug_enable(True)
    return h, th, tt, data


@pytest.mark.parametrize("shape,A,B,solver", [((1, 20, 20), 4, 8, G.SOLVER_ADAM), ((3, 28, 36), 3, 5, G.SOLVER_RMSPROP),
                                              ((1, 84, 84), 6, 20, G.SOLVER_RMSPROP)])   # the last = gold's Atari config
def test_conv_policy_learn_step(shape, A, B, solver):
    C, H, W = shape
    N = 64
    lr = 1e-3 if solver == G.SOLVER_RMSPROP else 5e-4
    h, th, tt, data = _make(shape, A, B, N, solver=solver, lr=lr)
    P = h.P
    w, m, v = th.copy(), np.zeros(P, np.float32), np.zeros(P, np.float32)
    w64, m64, v64 = th.astype(np.float64), np.zeros(P), np.zeros(P)
    it = 0
    for step in range(2):
        idx = synth.sample_indices(N, B, step)
        s, a, r, s2, d = O.gather(C * H * W, *data, idx)
        w_prev, v_prev = w.copy(), v.copy()
        out = O.atari_learn(shape, A, B, w, tt, m, v, it, s, a, r, s2, d, lr=lr, solver=solver)
        out64 = O.atari_learn(shape, A, B, w64, tt.astype(np.float64), m64, v64, it, s, a, r, s2, d, lr=lr, solver=solver,
                              dtype=np.float64)
        it = out["iter"]
        h.learn(idx)
        np.testing.assert_array_equal(h.debug_fetch(G.DBG_Q_ONLINE), out["q_online"])
        qt = h.debug_fetch(G.DBG_Q_TARGET)
        np.testing.assert_array_equal(qt[d == 0], out["q_target"][d == 0])
        np.testing.assert_array_equal(h.debug_fetch(G.DBG_TD), out["td"])
        loss = float(h.last_loss()[0])
        assert abs(loss - out64["loss"]) <= 1e-5 * abs(out64["loss"]), (loss, out64["loss"])
        g = h.debug_fetch(G.DBG_GRADS)
        gs = np.abs(out64["grads"]).max()
        np.testing.assert_allclose(g, out64["grads"], rtol=1e-4, atol=2e-6 * gs)
        w_gpu = h.get_params(G.NET_ONLINE)
        if solver == G.SOLVER_ADAM:
            frac = assert_params_close(w_gpu, w64, out64["grads"], lr=lr)
            assert frac < 0.05
        else:
            # RMSProp's step eta*g/sqrt((1-rho)*g^2 + eps) is ill-conditioned around |g| ~ sqrt(eps/(1-rho)) whatever the
            # largest gradient is, so the solver is checked exactly on the GPU's OWN gradient (op-for-op float32 replay
            # of solvers.go:273-335) and against the fp64 oracle by the bound of one step
            f = np.float32
            cache = (v_prev * f(0.999) + (g * g) * f(1.0 - 0.999)).astype(np.float32)
            want = (w_prev + (f(1) / np.sqrt(cache + f(1e-8))).astype(np.float32) * (g * f(-lr))).astype(np.float32)
            np.testing.assert_allclose(w_gpu, want, rtol=1e-6, atol=1e-8)
            assert np.abs(w_gpu - w64).max() <= 2 * lr / np.sqrt(1.0 - 0.999)
            assert np.mean(np.abs(w_gpu - w64) <= 1e-5 * np.abs(w64) + 1e-6) > 0.9
        # lock step: adopt the oracle's fp32 state
        h.set_params(G.NET_ONLINE, w); h.set_adam_state(m, v, it)
        w64[:] = w; m64[:] = m; v64[:] = v
    h.close()


def test_conv_policy_predict_act_and_graphs():
    shape, A, B, N = (1, 20, 20), 4, 8, 64
    C, H, W = shape
    h, th, tt, data = _make(shape, A, B, N)
    x = synth.replay((C * H * W, 0, 0, A), 50, seed=5)[0]
    q = O.atari_predict(shape, A, th, x)
    np.testing.assert_array_equal(h.predict(G.NET_ONLINE, x), q)
    np.testing.assert_array_equal(h.predict(G.NET_TARGET, x), O.atari_predict(shape, A, tt, x))
    np.testing.assert_array_equal(h.greedy_action(x), np.array([O.argmax(row) for row in q], np.int32))
    # learn_many (graphs) == the same steps one by one, and target sync copies all P parameters
    h2, *_ = _make(shape, A, B, N)
    h.learn_many(5, 100)
    for i in range(5):
        h2.learn(None, 100 + i)
    np.testing.assert_array_equal(h.get_params(G.NET_ONLINE), h2.get_params(G.NET_ONLINE))
    h.sync_target()
    np.testing.assert_array_equal(h.get_params(G.NET_TARGET), h.get_params(G.NET_ONLINE))
    with pytest.raises(G.GoldqError):
        G.Handle(arch=G.ARCH_ATARI_CONV, in_channels=1, in_height=84, in_width=84, obs_dim=100, n_actions=4)   # shape mismatch
    with pytest.raises(G.GoldqError):
        G.Handle(arch=G.ARCH_ATARI_CONV, in_channels=1, in_height=20, in_width=20, obs_dim=400, n_actions=4,
                 precision=G.PREC_BF16)                                                                       # no silent fallback
    h.close(); h2.close()
