"""The oracle's restatement of gold's Atari (convolutional) policy (oracle/conv_core.inc) against an independent
implementation: torch conv2d + autograd in float64.  Forward Q, loss and every gradient block must agree to
1e-10; the fp32 op-for-op version must sit within fp32 noise of its fp64 twin.  (CPU only.)"""
import numpy as np
import pytest
import torch

from oracle import binding as O


def _split(theta, C, F, A):
    sizes = [(32, C, 8, 8), (64, 32, 4, 4), (64, 64, 3, 3), (F, 512), (512,), (512, A), (A,)]
    out, o = [], 0
    for sh in sizes:
        n = int(np.prod(sh))
        out.append(torch.tensor(theta[o:o + n].reshape(sh), dtype=torch.float64, requires_grad=True))
        o += n
    assert o == theta.size
    return out


def _torch_q(params, x):
    f1, f2, f3, w4, b4, w5, b5 = params
    y = torch.relu(torch.nn.functional.conv2d(x, f1, stride=4, padding=1))
    y = torch.relu(torch.nn.functional.conv2d(y, f2, stride=2, padding=1))
    y = torch.relu(torch.nn.functional.conv2d(y, f3, stride=1, padding=1))
    y = torch.relu(y.reshape(x.shape[0], -1) @ w4 + b4)
    return y @ w5 + b5


@pytest.mark.parametrize("shape,A,B", [((1, 20, 20), 4, 6), ((3, 28, 36), 3, 4)])
def test_conv_policy_learn_step_matches_torch_autograd(shape, A, B):
    C, H, W = shape
    P, F = O.atari_nparams(C, H, W, A), O.atari_features(C, H, W)
    rng = np.random.default_rng(7)
    th = rng.standard_normal(P) * 0.05
    tt = rng.standard_normal(P) * 0.05
    s, s2 = rng.uniform(-1, 1, (B, C, H, W)), rng.uniform(-1, 1, (B, C, H, W))
    a = rng.integers(0, A, B).astype(np.int32)
    r = rng.uniform(0, 1, B)
    done = (rng.uniform(0, 1, B) < 0.3).astype(np.uint8)
    w = th.copy(); m = np.zeros(P); v = np.zeros(P)
    out = O.atari_learn(shape, A, B, w, tt, m, v, 0, s, a, r, s2, done, gamma=0.875, dtype=np.float64)
    # independent restatement
    po = _split(th, C, F, A)
    with torch.no_grad():
        qt = _torch_q(_split(tt, C, F, A), torch.tensor(s2))
        td = torch.tensor(r) + 0.875 * qt.max(dim=1).values * torch.tensor(1.0 - done)
    q = _torch_q(po, torch.tensor(s))
    y = q.detach().clone()
    y[torch.arange(B), torch.tensor(a, dtype=torch.long)] = td
    loss = ((q - y) ** 2).mean()
    loss.backward()
    np.testing.assert_allclose(out["q_online"], q.detach().numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(out["td"], td.numpy(), rtol=1e-10, atol=1e-12)
    assert abs(out["loss"] - float(loss)) <= 1e-10 * abs(float(loss))
    g = np.concatenate([p.grad.numpy().ravel() for p in po])
    np.testing.assert_allclose(out["grads"], g, rtol=1e-9, atol=1e-12 * np.abs(g).max())
    # the fp32 op-for-op twin is within fp32 noise
    w32 = th.astype(np.float32); m32 = np.zeros(P, np.float32); v32 = np.zeros(P, np.float32)
    o32 = O.atari_learn(shape, A, B, w32, tt.astype(np.float32), m32, v32, 0, s, a, r, s2, done, gamma=0.875)
    np.testing.assert_allclose(o32["q_online"], out["q_online"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(o32["grads"], out["grads"], rtol=1e-3, atol=1e-5 * np.abs(g).max())


def test_conv_policy_shapes_are_golds():
    """84 x 84 inputs give the 6400 features policy.go:68 hard-codes for the first FC layer."""
    assert O.atari_features(1, 84, 84) == 6400
    assert O.atari_nparams(1, 84, 84, 6) == 32 * 64 + 64 * 32 * 16 + 64 * 64 * 9 + 6400 * 512 + 512 + 512 * 6 + 6
