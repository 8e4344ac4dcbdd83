"""Known-answer tests that pin the oracle.

Reference material (all the reference's own tests hold for this path):
  pkg/v1/dense/op_test.go:12-18   AMax {5,8,3,2} -> 8
  pkg/v1/dense/op_test.go:20-31   Concat of three (1,5) rows -> (3,5)
  pkg/v1/agent/deepq/memory_test.go:12-29  push 10 events, Sample(5) -> 5 events
plus closed-form answers derived from the reference source (cited per test).
"""
import numpy as np

from gold_b200 import synth
from oracle import binding as O


def test_amax_reference_kat():
    # pkg/v1/dense/op_test.go:12-18
    assert O.amax(np.array([5, 8, 3, 2], np.float32)) == 8.0
    assert O.argmax(np.array([5, 8, 3, 2], np.float32)) == 1


def test_argmax_first_max_and_nan_inf_rules():
    # vendor/gorgonia.org/tensor/internal/execution/generic_argmethods.go:211-233
    assert O.argmax(np.array([1, 3, 3, 2], np.float32)) == 1          # ties -> first
    assert O.argmax(np.array([7, 7], np.float32)) == 0
    assert O.argmax(np.array([9, np.nan, 10], np.float32)) == 1       # NaN at i>=1 wins at once
    assert O.argmax(np.array([9, 1, np.inf, 10], np.float32)) == 2    # +Inf at i>=1 wins at once
    assert O.argmax(np.array([np.nan, 1, 2], np.float32)) == 0        # index 0 not special-cased: v > NaN never true
    assert O.argmax(np.array([-np.inf, -np.inf], np.float32)) == 0


def test_concat_and_sample_length_kat():
    # op_test.go:20-31: Concat(0, three (1,5) tensors) has shape (3,5);
    # memory_test.go:28: Sample(5) of 10 events returns 5.
    D = 5
    s = np.arange(10 * D, dtype=np.float32).reshape(10, D)
    z = np.zeros(10, np.float32)
    idx = synth.sample_indices(10, 5, 0)
    assert len(idx) == 5 and len(set(idx.tolist())) == 5 and idx.min() >= 0 and idx.max() < 10
    gs, ga, gr, gs2, gd = O.gather(D, s, z.astype(np.int32), z, s, z.astype(np.uint8), idx[:3])
    assert gs.shape == (3, 5)
    # logical index i = i-th newest (PushFront / At): newest row is arrival row 9
    for b, i in enumerate(idx[:3]):
        np.testing.assert_array_equal(gs[b], s[9 - i])


def test_forward_closed_form():
    # y = relu(relu(x W1 + b1) W2 + b2) W3 + b3 with W (in x out) row-major
    dims = (2, 2, 2, 2)
    th = np.array([1, -1, 2, 0.5,   0.5, 0.25,        # W1 (2x2), b1
                   1, 0, 0, 1,      0, 0,             # W2 = I, b2
                   1, 2, 3, 4,      0.5, -0.5],       # W3, b3
                  np.float32)
    x = np.array([[1.0, 1.0]], np.float32)
    # a1 = [1*1+1*2+0.5, 1*-1+1*0.5+0.25] = [3.5, -0.25]; h1 = [3.5, -0.0]
    # h2 = h1; q = [3.5*1 + (-0)*3 + 0.5, 3.5*2 + (-0)*4 - 0.5] = [4.0, 6.5]
    q, g = O.predict(dims, th, x)
    np.testing.assert_array_equal(q, np.array([[4.0, 6.5]], np.float32))
    assert g[0] == 1


def test_relu_negative_zero_and_mask_at_zero():
    # vendor/gorgonia.org/gorgonia/nn.go:162-189: h = a * (a >= 0 ? 1 : 0)
    dims = (1, 1, 1, 1)
    th = np.array([1, 0,  1, 0,  1, 0], np.float32)
    q, _ = O.predict(dims, th, np.array([[-3.0]], np.float32))
    # h1 = -3*0 = -0.0, but the next Sgemm skips terms with x_k == 0 (sgemm.go:263-267) and
    # +0 + (-0) = +0 anyway, so the sign never reaches the output
    assert q[0, 0] == 0.0 and not np.signbit(q[0, 0])
    # derivative at exactly 0 is 1: gradient flows through a unit whose pre-activation is 0
    th = np.array([0, 0,  1, 0,  1, 0], np.float32)       # a1 = 0 for any x
    m = np.zeros(6, np.float32); v = np.zeros(6, np.float32)
    out = O.learn(dims, 1, th.copy(), th.copy(), m, v, 0,
                  np.array([[2.0]], np.float32), [0], [1.0], np.array([[2.0]], np.float32), [1])
    # q = 0, y = 1 -> dq = 2*(0-1)/1 = -2 ; dW3 = h2*dq = 0 ; db3 = -2 ; db2 = -2*mask(0)=−2 ; db1 = -2 ; dW1 = x*db1 = -4
    np.testing.assert_array_equal(out["grads"], np.array([-4, -2, 0, -2, 0, -2], np.float32))


def test_adam_step1_closed_form_double_update():
    # vendor/gorgonia.org/gorgonia/solvers.go:440-617: after step 1,
    # m = (1-b1) g, v = (1-b2) g^2, mhat = g, vhat = g^2,
    # w1 = w0 - eta*g/(|g|+eps) - eta*g   (Div WithIncr(w) at :604, then Add(w, mHats) at :611)
    g = np.array([0.5, -2.0, 1e-3, 0.0], np.float32)
    w = np.array([1.0, 1.0, 1.0, 1.0], np.float32)
    m = np.zeros(4, np.float32)
    v = np.zeros(4, np.float32)
    g_in = g.copy()
    O.adam(w, g_in, m, v, 1, lr=5e-4)
    eta = 5e-4
    expect = 1.0 - eta * np.sign(g) - eta * g
    np.testing.assert_allclose(w, expect, rtol=0, atol=2e-7)
    assert w[3] == 1.0
    np.testing.assert_array_equal(g_in, 0)                      # g.Zero()
    # omb = float32(1) - float32(beta) (solvers.go:497-498): 1-0.999 carries cancellation error
    omb1 = np.float32(1) - np.float32(0.9)
    omb2 = np.float32(1) - np.float32(0.999)
    np.testing.assert_array_equal(m, g * omb1)
    np.testing.assert_array_equal(v, (g * g) * omb2)
    # a textbook single-update Adam would give 1 - eta*sign(g): must differ for g = -2
    assert abs(w[1] - (1.0 + eta)) > 5e-4


def test_td_target_and_label_row():
    # agent.go:137-155: done -> target = r ; else r + gamma*max_a Q_target(s')
    dims = (1, 1, 1, 2)
    # online: q = [h, 2h] with h = relu(relu(x)) ; target: q = [3h, h]
    th_on = np.array([1, 0, 1, 0, 1, 2, 0, 0], np.float32)
    th_tg = np.array([1, 0, 1, 0, 3, 1, 0, 0], np.float32)
    P = 8
    out = O.learn(dims, 2, th_on.copy(), th_tg, np.zeros(P, np.float32), np.zeros(P, np.float32), 0,
                  np.array([[1.0], [2.0]], np.float32), [1, 0], [0.5, 0.25],
                  np.array([[2.0], [4.0]], np.float32), [0, 1], gamma=0.5)
    np.testing.assert_array_equal(out["q_online"], np.array([[1, 2], [2, 4]], np.float32))
    np.testing.assert_array_equal(out["q_target"][0], np.array([6, 2], np.float32))
    np.testing.assert_array_equal(out["td"], np.array([0.5 + 0.5 * 6, 0.25], np.float32))
    # loss = mean over B*A of (q - y)^2 with y = q except the action entry
    expect = ((2 - 3.5) ** 2 + (2 - 0.25) ** 2) / 4
    assert abs(out["loss"] - expect) < 1e-7
    assert out["iter"] == 1


def test_threads_do_not_change_results():
    dims = synth.CARTPOLE
    B = 256
    th = synth.glorot_params(dims, 1)
    tt = synth.glorot_params(dims, 2)
    s, a, r, s2, d = synth.replay(dims, B)
    P = synth.nparams(dims)
    res = []
    for t in (1, 4):
        O.set_threads(t)
        w = th.copy(); m = np.zeros(P, np.float32); v = np.zeros(P, np.float32)
        out = O.learn(dims, B, w, tt, m, v, 0, s, a, r, s2, d)
        res.append((w, m, v, out["grads"], out["loss"]))
    O.set_threads(1)
    for x, y in zip(res[0], res[1]):
        np.testing.assert_array_equal(x, y)
