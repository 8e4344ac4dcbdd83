"""Independent cross-check of the oracle: torch autograd (CPU, float64) computes
the same loss / gradients from the mathematical definition of the learn step
(SURVEY.md Appendix A), without sharing any code with oracle/.  This does not
replace executed reference output (parity stays unpinned, see oracle header)
but it catches transcription errors in the hand-derived backward pass."""
import numpy as np
import pytest
import torch

from gold_b200 import synth
from oracle import binding as O
from util import assert_params_close


def _torch_step(dims, th, tt, s, a, r, s2, done, gamma, huber_delta=None):
    D, H1, H2, A = dims
    off = synth.offsets(dims)

    def unpack(t):
        g = lambda k, shape: t[off[k][0]:off[k][1]].reshape(shape)
        return (g("w1", (D, H1)), g("b1", (H1,)), g("w2", (H1, H2)), g("b2", (H2,)),
                g("w3", (H2, A)), g("b3", (A,)))

    def fwd(t, x):
        w1, b1, w2, b2, w3, b3 = unpack(t)
        h1 = torch.relu(x @ w1 + b1)
        h2 = torch.relu(h1 @ w2 + b2)
        return h2 @ w3 + b3

    th_t = torch.tensor(th, dtype=torch.float64, requires_grad=True)
    tt_t = torch.tensor(tt, dtype=torch.float64)
    S = torch.tensor(s, dtype=torch.float64)
    S2 = torch.tensor(s2, dtype=torch.float64)
    R = torch.tensor(r, dtype=torch.float64)
    with torch.no_grad():
        qn = fwd(tt_t, S2).max(dim=1).values
        td = torch.where(torch.tensor(done.astype(bool)), R, R + gamma * qn)
        y = fwd(th_t, S).clone()
        y[torch.arange(len(a)), torch.tensor(a, dtype=torch.long)] = td
    q = fwd(th_t, S)
    if huber_delta is None:
        loss = ((q - y) ** 2).mean()
    else:           # pseudo-Huber as goro defines it (loss.go:113-145), reduced by the mean
        loss = (huber_delta ** 2 * (torch.sqrt(1 + ((q - y) / huber_delta) ** 2) - 1)).mean()
    loss.backward()
    return q.detach().numpy(), td.numpy(), loss.item(), th_t.grad.numpy()


@pytest.mark.parametrize("dims,B", [(synth.CARTPOLE, 20), (synth.CARTPOLE, 512), ((16, 48, 32, 6), 96)])
def test_oracle_f64_matches_autograd(dims, B):
    th = synth.glorot_params(dims, 1).astype(np.float64)
    tt = synth.glorot_params(dims, 2).astype(np.float64)
    s, a, r, s2, done = synth.replay(dims, B, p_done=0.2)
    P = synth.nparams(dims)
    w = th.copy()
    out = O.learn(dims, B, w, tt, np.zeros(P), np.zeros(P), 0, s.astype(np.float64), a,
                  r.astype(np.float64), s2.astype(np.float64), done, gamma=0.95, dtype=np.float64)
    q, td, loss, grads = _torch_step(dims, th, tt, s, a, r, s2, done, float(np.float32(0.95)))  # Gamma is a float32 field (agent.go:48)
    np.testing.assert_allclose(out["q_online"], q, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(out["td"], td, rtol=1e-12, atol=1e-13)
    assert abs(out["loss"] - loss) <= 1e-12 * abs(loss)
    np.testing.assert_allclose(out["grads"], grads, rtol=1e-10, atol=1e-14)


@pytest.mark.parametrize("dims,B", [(synth.CARTPOLE, 20), (synth.CARTPOLE, 4096), (synth.WIDE, 256)])
def test_oracle_f32_within_budget_of_f64(dims, B):
    """The fp32 restatement must sit within fp32 rounding of its fp64 twin."""
    th = synth.glorot_params(dims, 1)
    tt = synth.glorot_params(dims, 2)
    s, a, r, s2, done = synth.replay(dims, B)
    P = synth.nparams(dims)
    w32 = th.copy(); m32 = np.zeros(P, np.float32); v32 = np.zeros(P, np.float32)
    o32 = O.learn(dims, B, w32, tt, m32, v32, 0, s, a, r, s2, done)
    w64 = th.astype(np.float64); m64 = np.zeros(P); v64 = np.zeros(P)
    o64 = O.learn(dims, B, w64, tt.astype(np.float64), m64, v64, 0, s.astype(np.float64), a,
                  r.astype(np.float64), s2.astype(np.float64), done, dtype=np.float64)
    np.testing.assert_allclose(o32["q_online"], o64["q_online"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(o32["td"], o64["td"], rtol=2e-5, atol=2e-6)
    assert abs(o32["loss"] - o64["loss"]) <= 2e-5 * abs(o64["loss"])
    gscale = np.abs(o64["grads"]).max()
    np.testing.assert_allclose(o32["grads"], o64["grads"], rtol=1e-4, atol=1e-5 * gscale)
    frac = assert_params_close(w32, w64, o64["grads"])
    assert frac < 0.01


@pytest.mark.parametrize("delta", [1.0, 0.25])
def test_oracle_pseudo_huber_matches_autograd(delta):
    dims, B = (16, 48, 32, 6), 96
    th = synth.glorot_params(dims, 1).astype(np.float64)
    tt = synth.glorot_params(dims, 2).astype(np.float64)
    s, a, r, s2, done = synth.replay(dims, B, p_done=0.2)
    P = synth.nparams(dims)
    out = O.learn(dims, B, th.copy(), tt, np.zeros(P), np.zeros(P), 0, s.astype(np.float64), a, r.astype(np.float64),
                  s2.astype(np.float64), done, gamma=0.95, dtype=np.float64, loss=1, delta=delta)
    _, _, loss, grads = _torch_step(dims, th, tt, s, a, r, s2, done, float(np.float32(0.95)), huber_delta=delta)
    assert abs(out["loss"] - loss) <= 1e-12 * abs(loss)
    np.testing.assert_allclose(out["grads"], grads, rtol=1e-10, atol=1e-14)


def test_oracle_rmsprop_is_the_reference_op_sequence():
    """RMSPropSolver.Step, float32 branch (gorgonia solvers.go:273-335), replayed with numpy float32
    scalars in the same order: Square, cw*=decay, g2*=omdecay, cw+=g2, +eps, 1/sqrt, g*=-eta, *, w+=."""
    dims, B = synth.CARTPOLE, 20
    th = synth.glorot_params(dims, 1); tt = synth.glorot_params(dims, 2)
    s, a, r, s2, done = synth.replay(dims, B)
    P = synth.nparams(dims)
    f = np.float32
    eta, rho, eps = 1e-3, 0.999, 1e-8
    decay, omdecay, e32, step = f(rho), f(1.0 - rho), f(eps), f(-eta)
    w = th.copy(); m = np.zeros(P, f); cache = np.zeros(P, f)
    w_ref, c_ref = th.copy(), np.zeros(P, f)
    for it in range(3):
        out = O.learn(dims, B, w, tt, m, cache, it, s, a, r, s2, done, lr=eta, beta2=rho, eps=eps, solver=1)
        g = out["grads"].astype(f)
        c_ref = (c_ref * decay + (g * g) * omdecay).astype(f)
        upd = (f(1) / np.sqrt((c_ref + e32).astype(f), dtype=f)).astype(f)
        w_ref = (w_ref + upd * (g * step).astype(f)).astype(f)
        np.testing.assert_array_equal(cache, c_ref)
        np.testing.assert_array_equal(w, w_ref)
    assert not m.any()                        # RMSProp keeps one cache; the Adam first moment stays untouched
