"""ctypes binding of the CPU oracle (oracle/goldq_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(gold_b200/) must never import this module.
"""
import ctypes as C
C_ = C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgoldq_oracle.so")


class OqDims(C.Structure):
    _fields_ = [("D", C.c_int32), ("H1", C.c_int32), ("H2", C.c_int32), ("A", C.c_int32)]


class OqCfg(C.Structure):
    _fields_ = [("dims", OqDims), ("B", C.c_int32), ("gamma", C.c_float),
                ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double), ("loss", C.c_int32), ("solver", C.c_int32), ("delta", C.c_float)]


def build(force=False):
    """Compile the oracle in place (gcc; a second or two)."""
    src = [os.path.join(_HERE, f) for f in ("goldq_oracle.c", "learn_core.inc", "conv_core.inc", "Makefile")]
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= max(os.path.getmtime(s) for s in src)):
        return _SO
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oq_nparams.restype = C.c_int64
        _lib.oq_amax_f32.restype = C.c_float
        _lib.oq_max_threads.restype = C.c_int
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def set_threads(n):
    lib().oq_set_threads(int(n))


def max_threads():
    return int(lib().oq_max_threads())


def nparams(D, H1, H2, A):
    return int(lib().oq_nparams(D, H1, H2, A))


def argmax(a):
    a = np.ascontiguousarray(a, np.float32)
    return int(lib().oq_argmax_f32(_p(a, C.c_float), a.size))


def amax(a):
    a = np.ascontiguousarray(a, np.float32)
    return float(lib().oq_amax_f32(_p(a, C.c_float), a.size))


def predict(dims, theta, x):
    """Returns (q [n,A] f32, greedy [n] i32)."""
    D, H1, H2, A = dims
    x = np.ascontiguousarray(x, np.float32).reshape(-1, D)
    theta = np.ascontiguousarray(theta, np.float32)
    n = x.shape[0]
    q = np.empty((n, A), np.float32)
    g = np.empty(n, np.int32)
    lib().oq_predict_f32(D, H1, H2, A, _p(theta, C.c_float), n, _p(x, C.c_float),
                         _p(q, C.c_float), _p(g, C.c_int32))
    return q, g


def predict64(dims, theta, x):
    D, H1, H2, A = dims
    x = np.ascontiguousarray(x, np.float64).reshape(-1, D)
    theta = np.ascontiguousarray(theta, np.float64)
    n = x.shape[0]
    q = np.empty((n, A), np.float64)
    lib().oq_predict_f64(D, H1, H2, A, _p(theta, C.c_double), n, _p(x, C.c_double),
                         _p(q, C.c_double))
    return q


def adam(w, g, m, v, it, lr=5e-4, beta1=0.9, beta2=0.999, eps=1e-8):
    """In-place on float32 arrays; `it` is the iteration number of THIS step (>=1)."""
    for a in (w, g, m, v):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    lib().oq_adam_f32(C.c_int64(w.size), _p(w, C.c_float), _p(g, C.c_float),
                      _p(m, C.c_float), _p(v, C.c_float), C.c_int64(it),
                      C.c_double(lr), C.c_double(beta1), C.c_double(beta2), C.c_double(eps))


def learn(dims, B, theta, theta_t, m, v, it, s, a, r, s2, done, gamma=0.95, lr=5e-4,
          beta1=0.9, beta2=0.999, eps=1e-8, dtype=np.float32, want_debug=True, loss=0, solver=0, delta=1.0):
    """One Agent.Learn on a sampled batch.  theta/m/v are updated IN PLACE.

    `it` = solver iteration count before the step.  Returns dict with the new
    `iter` and (if want_debug) q_online, q_target, td, loss, grads.
    """
    D, H1, H2, A = dims
    f32 = dtype == np.float32
    ct = C.c_float if f32 else C.c_double
    P = nparams(*dims)
    for arr in (theta, m, v):
        assert arr.dtype == dtype and arr.flags.c_contiguous and arr.size == P
    theta_t = np.ascontiguousarray(theta_t, dtype)
    s = np.ascontiguousarray(s, dtype).reshape(B, D)
    s2 = np.ascontiguousarray(s2, dtype).reshape(B, D)
    r = np.ascontiguousarray(r, dtype).reshape(B)
    a = np.ascontiguousarray(a, np.int32).reshape(B)
    done = np.ascontiguousarray(done, np.uint8).reshape(B)
    cfg = OqCfg(OqDims(D, H1, H2, A), B, gamma, lr, beta1, beta2, eps, loss, solver, delta)
    out = {}
    if want_debug:
        out["q_online"] = np.zeros((B, A), dtype)
        out["q_target"] = np.zeros((B, A), dtype)
        out["td"] = np.zeros(B, dtype)
        out["loss"] = np.zeros(1, dtype)
        out["grads"] = np.zeros(P, dtype)
    itc = C.c_int64(it)
    fn = lib().oq_learn_f32 if f32 else lib().oq_learn_f64
    rc = fn(C.byref(cfg), _p(theta, ct), _p(theta_t, ct), _p(m, ct), _p(v, ct),
            C.byref(itc), _p(s, ct), _p(a, C.c_int32), _p(r, ct), _p(s2, ct),
            _p(done, C.c_uint8), _p(out.get("q_online"), ct), _p(out.get("q_target"), ct),
            _p(out.get("td"), ct), _p(out.get("loss"), ct), _p(out.get("grads"), ct))
    assert rc == 0
    out["iter"] = itc.value
    if want_debug:
        out["loss"] = out["loss"][0]
    return out


def fit(dims, theta, m, v, it, x, y, lr=5e-4, beta1=0.9, beta2=0.999, eps=1e-8, dtype=np.float32, loss=0, solver=0,
        delta=1.0):
    """Sequential.FitBatch on explicit labels; theta/m/v updated IN PLACE."""
    D, H1, H2, A = dims
    f32 = dtype == np.float32
    ct = C.c_float if f32 else C.c_double
    P = nparams(*dims)
    x = np.ascontiguousarray(x, dtype).reshape(-1, D)
    y = np.ascontiguousarray(y, dtype).reshape(-1, A)
    n = x.shape[0]
    assert y.shape[0] == n
    for arr in (theta, m, v):
        assert arr.dtype == dtype and arr.flags.c_contiguous and arr.size == P
    cfg = OqCfg(OqDims(D, H1, H2, A), n, 0.0, lr, beta1, beta2, eps, loss, solver, delta)
    loss = np.zeros(1, dtype)
    grads = np.zeros(P, dtype)
    itc = C.c_int64(it)
    fn = lib().oq_fit_f32 if f32 else lib().oq_fit_f64
    rc = fn(C.byref(cfg), n, _p(theta, ct), _p(m, ct), _p(v, ct), C.byref(itc), _p(x, ct),
            _p(y, ct), _p(loss, ct), _p(grads, ct))
    assert rc == 0
    return {"iter": itc.value, "loss": loss[0], "grads": grads}


def gather(D, order_s, order_a, order_r, order_s2, order_done, idx):
    """Rows idx (i = i-th newest) out of arrival-ordered arrays."""
    n = order_a.shape[0]
    idx = np.ascontiguousarray(idx, np.int32)
    B = idx.size
    s = np.empty((B, D), np.float32)
    s2 = np.empty((B, D), np.float32)
    a = np.empty(B, np.int32)
    r = np.empty(B, np.float32)
    done = np.empty(B, np.uint8)
    order_s = np.ascontiguousarray(order_s, np.float32)
    order_s2 = np.ascontiguousarray(order_s2, np.float32)
    order_a = np.ascontiguousarray(order_a, np.int32)
    order_r = np.ascontiguousarray(order_r, np.float32)
    order_done = np.ascontiguousarray(order_done, np.uint8)
    lib().oq_gather_f32(D, C.c_int64(n), _p(order_s, C.c_float), _p(order_a, C.c_int32),
                        _p(order_r, C.c_float), _p(order_s2, C.c_float),
                        _p(order_done, C.c_uint8), B, _p(idx, C.c_int32), _p(s, C.c_float),
                        _p(a, C.c_int32), _p(r, C.c_float), _p(s2, C.c_float),
                        _p(done, C.c_uint8))
    return s, a, r, s2, done


# ---- the Atari (convolutional) policy (conv_core.inc) ------------------------------------------------------
def atari_nparams(C, H, W, A):
    lib().oq_atari_nparams.restype = C_.c_int64
    return int(lib().oq_atari_nparams(C, H, W, A))


def atari_features(C, H, W):
    return int(lib().oq_atari_features(C, H, W))


def atari_predict(shape, A, theta, x, dtype=np.float32):
    """q [n][A] of the conv policy for n states [n][C][H][W]."""
    C, H, W = shape
    ct = C_.c_float if dtype == np.float32 else C_.c_double
    theta = np.ascontiguousarray(theta, dtype)
    x = np.ascontiguousarray(x, dtype).reshape(-1, C * H * W)
    q = np.empty((x.shape[0], A), dtype)
    fn = lib().oq_atari_predict_f32 if dtype == np.float32 else lib().oq_atari_predict_f64
    fn(C, H, W, A, _p(theta, ct), x.shape[0], _p(x, ct), _p(q, ct))
    return q


def atari_learn(shape, A, B, theta, theta_t, m, v, it, s, a, r, s2, done, gamma=0.95, lr=5e-4, beta1=0.9, beta2=0.999,
                eps=1e-8, dtype=np.float32, solver=0):
    """One Agent.Learn of the conv policy on a sampled batch; theta/m/v updated IN PLACE."""
    C, H, W = shape
    f32 = dtype == np.float32
    ct = C_.c_float if f32 else C_.c_double
    P = atari_nparams(C, H, W, A)
    for arr in (theta, m, v):
        assert arr.dtype == dtype and arr.flags.c_contiguous and arr.size == P
    theta_t = np.ascontiguousarray(theta_t, dtype)
    s = np.ascontiguousarray(s, dtype).reshape(B, -1)
    s2 = np.ascontiguousarray(s2, dtype).reshape(B, -1)
    r = np.ascontiguousarray(r, dtype).reshape(B)
    a = np.ascontiguousarray(a, np.int32).reshape(B)
    done = np.ascontiguousarray(done, np.uint8).reshape(B)
    cfg = OqCfg(OqDims(C * H * W, 0, 0, A), B, gamma, lr, beta1, beta2, eps, 0, solver, 1.0)
    out = {"q_online": np.zeros((B, A), dtype), "q_target": np.zeros((B, A), dtype), "td": np.zeros(B, dtype),
           "loss": np.zeros(1, dtype), "grads": np.zeros(P, dtype)}
    itc = C_.c_int64(it)
    fn = lib().oq_atari_learn_f32 if f32 else lib().oq_atari_learn_f64
    rc = fn(C_.byref(cfg), C, H, W, _p(theta, ct), _p(theta_t, ct), _p(m, ct), _p(v, ct), C_.byref(itc), _p(s, ct),
            _p(a, C_.c_int32), _p(r, ct), _p(s2, ct), _p(done, C_.c_uint8), _p(out["q_online"], ct), _p(out["q_target"], ct),
            _p(out["td"], ct), _p(out["loss"], ct), _p(out["grads"], ct))
    assert rc == 0
    out["iter"] = itc.value
    out["loss"] = out["loss"][0]
    return out
