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
