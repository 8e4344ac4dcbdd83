"""Shared comparison rules."""
import numpy as np


def assert_params_close(w, w_ref, g_ref, lr=5e-4, rtol=1e-5, atol=1e-7, cond=1e-5, what="params"):
    """Post-Adam parameters.

    Adam's early steps are sign-like: u = -lr*g/(|g|+eps) has sensitivity
    lr*eps/(|g|+eps)^2 to gradient noise, unbounded in relative terms as g -> 0.
    Elements whose reference gradient is ill-conditioned (|g| < cond*max|g|) are
    therefore 'ties' (same spirit as the argmax-tie exclusion): for them only the
    hard bound |dw| <= 2*lr*(1+|g|) is asserted.  Everything else must meet
    rtol/atol.  The fraction of excluded elements is asserted small.
    """
    w = np.asarray(w, np.float64)
    w_ref = np.asarray(w_ref, np.float64)
    g_ref = np.abs(np.asarray(g_ref, np.float64))
    # exact zeros (dead ReLU units) are well-conditioned: the weight must not move
    well = (g_ref >= cond * g_ref.max()) | (g_ref == 0)
    err = np.abs(w - w_ref)
    tol = rtol * np.maximum(np.abs(w), np.abs(w_ref)) + atol
    bad = well & (err > tol)
    assert not bad.any(), (f"{what}: {bad.sum()} well-conditioned elements out of tolerance; "
                           f"worst abs err {err[bad].max():.3e}")
    ill = ~well
    if ill.any():
        assert (err[ill] <= 2 * lr * (1 + g_ref[ill]) + tol[ill]).all(), f"{what}: ill-conditioned element beyond 2*lr bound"
    return float(ill.mean())


def rel_err(x, ref):
    x = np.asarray(x, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-300))


class OracleAgent:
    """Oracle-side twin of one goldq handle (test infrastructure).

    Holds the reference-arithmetic state (fp32) and an fp64 shadow used only to
    budget the error of order-dependent sums (loss, gradients)."""

    def __init__(self, dims, B, theta, theta_t, gamma=0.95, lr=5e-4, capacity=None, loss=0, solver=0, delta=1.0,
                 beta2=0.999, eps=1e-8):
        from oracle import binding as O
        self.O = O
        self.dims, self.B, self.gamma, self.lr = dims, B, gamma, lr
        self.opt = dict(loss=loss, solver=solver, delta=delta, beta2=beta2, eps=eps)
        # bound on one solver step for ill-conditioned elements: Adam moves <= 2*lr, RMSProp's first steps
        # are -lr*sign(g)/sqrt(1-rho)
        self.bound_lr = lr if solver == 0 else lr / np.sqrt(1.0 - beta2)
        P = O.nparams(*dims)
        self.theta = np.array(theta, np.float32).copy()
        self.theta_t = np.array(theta_t, np.float32).copy()
        self.m = np.zeros(P, np.float32)
        self.v = np.zeros(P, np.float32)
        self.iter = 0
        self.capacity = capacity
        D = dims[0]
        self.s = np.zeros((0, D), np.float32)
        self.s2 = np.zeros((0, D), np.float32)
        self.a = np.zeros(0, np.int32)
        self.r = np.zeros(0, np.float32)
        self.done = np.zeros(0, np.uint8)

    def push(self, s, a, r, s2, done):
        self.s = np.concatenate([self.s, s]); self.s2 = np.concatenate([self.s2, s2])
        self.a = np.concatenate([self.a, a]); self.r = np.concatenate([self.r, r])
        self.done = np.concatenate([self.done, done])
        if self.capacity is not None and len(self.a) > self.capacity:   # PopBack: drop the oldest
            k = len(self.a) - self.capacity
            self.s, self.s2, self.a, self.r, self.done = (x[k:] for x in (self.s, self.s2, self.a, self.r, self.done))

    def batch(self, idx):
        return self.O.gather(self.dims[0], self.s, self.a, self.r, self.s2, self.done, idx)

    def learn(self, idx, with64=True):
        s, a, r, s2, done = self.batch(idx)
        out64 = None
        if with64:
            th64 = self.theta.astype(np.float64); m64 = self.m.astype(np.float64); v64 = self.v.astype(np.float64)
            out64 = self.O.learn(self.dims, self.B, th64, self.theta_t.astype(np.float64), m64, v64, self.iter,
                                 s.astype(np.float64), a, r.astype(np.float64), s2.astype(np.float64), done,
                                 gamma=self.gamma, lr=self.lr, dtype=np.float64, **self.opt)
            out64["theta"] = th64
        out = self.O.learn(self.dims, self.B, self.theta, self.theta_t, self.m, self.v, self.iter,
                           s, a, r, s2, done, gamma=self.gamma, lr=self.lr, **self.opt)
        self.iter = out["iter"]
        out["done"] = done
        return out, out64

    def sync_target(self):
        self.theta_t = self.theta.copy()
