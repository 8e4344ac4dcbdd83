"""ctypes binding of include/goldq.h (libgoldq.so).

There is deliberately no fallback: if the CUDA library is missing or a call
fails, an exception is raised.  Nothing here imports oracle/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GOLDQ_LIB: another build of the same library (diagnostics: lib/libgoldq_trace.so from `make trace`)
LIB_PATH = os.environ.get("GOLDQ_LIB") or os.path.join(_HERE, "lib", "libgoldq.so")

OK, EINVAL, ECUDA, ENOTREADY, EUNSUPPORTED, ENCCL = 0, -1, -2, -3, -4, -5
NET_ONLINE, NET_TARGET = 0, 1
LOSS_MSE, LOSS_PSEUDO_HUBER = 0, 1
SOLVER_ADAM, SOLVER_RMSPROP = 0, 1
PREC_FP32, PREC_BF16 = 0, 1
PATH_AUTO, PATH_GENERIC, PATH_NARROW, PATH_WIDE, PATH_CONV = 0, 1, 2, 3, 4
ARCH_MLP, ARCH_ATARI_CONV = 0, 1
DBG_Q_ONLINE, DBG_Q_TARGET, DBG_TD, DBG_GRADS, DBG_INDICES, DBG_BATCH_S, DBG_BATCH_A, DBG_FF_STAMPS, DBG_TRACE = range(9)
ABI_VERSION = 2

# every symbol include/goldq.h declares (tests/test_abi.py checks header <-> list <-> .so)
SYMBOLS = [
    "goldq_default_config", "goldq_create", "goldq_destroy", "goldq_last_error",
    "goldq_abi_version", "goldq_param_count", "goldq_selected_path", "goldq_set_params",
    "goldq_get_params", "goldq_set_adam_state", "goldq_get_adam_state", "goldq_replay_push",
    "goldq_replay_len", "goldq_replay_push_wait", "goldq_replay_stage_packed", "goldq_replay_commit", "goldq_host_alloc", "goldq_host_free",
    "goldq_replay_push_packed", "goldq_packed_bytes", "goldq_learn_prepare", "goldq_graph_builds", "goldq_launch_floor", "goldq_time_forward", "goldq_resize_batch", "goldq_batch_size",
    "goldq_learn", "goldq_learn_device_indices", "goldq_learn_many", "goldq_sampler_indices", "goldq_sync_target",
    "goldq_last_loss", "goldq_predict", "goldq_greedy_action", "goldq_epsilon_greedy_action", "goldq_explore_plan",
    "goldq_fit_batch",
    "goldq_dp_unique_id", "goldq_dp_init", "goldq_dp_mode", "goldq_sync", "goldq_stream", "goldq_launch_count",
    "goldq_debug_enable", "goldq_debug_fetch", "goldq_timer_start", "goldq_timer_stop",
    "goldq_selftest_umma", "goldq_step_breakdown",
]


class GoldqConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("obs_dim", C.c_int32), ("hidden1", C.c_int32),
        ("hidden2", C.c_int32), ("n_actions", C.c_int32), ("batch_size", C.c_int32),
        ("global_batch", C.c_int32), ("replay_capacity", C.c_int64), ("gamma", C.c_float), ("huber_delta", C.c_float),
        ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
        ("loss", C.c_int32), ("precision", C.c_int32), ("path", C.c_int32), ("device", C.c_int32),
        ("n_replicas", C.c_int32), ("solver", C.c_int32), ("stream", C.c_void_p),
        ("arch", C.c_int32), ("in_channels", C.c_int32), ("in_height", C.c_int32), ("in_width", C.c_int32),
    ]


class GoldqError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"goldq error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libgoldq.so (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C gold_b200/csrc).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    fp, ip, up = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
    sig = {
        "goldq_default_config": (C.c_int, [C.POINTER(GoldqConfig)]),
        "goldq_create": (vp, [C.POINTER(GoldqConfig)]),
        "goldq_destroy": (None, [vp]),
        "goldq_last_error": (C.c_char_p, [vp]),
        "goldq_abi_version": (C.c_int, []),
        "goldq_param_count": (i64, [vp]),
        "goldq_selected_path": (C.c_int, [vp]),
        "goldq_set_params": (C.c_int, [vp, C.c_int, fp]),
        "goldq_get_params": (C.c_int, [vp, C.c_int, fp]),
        "goldq_set_adam_state": (C.c_int, [vp, fp, fp, i64]),
        "goldq_get_adam_state": (C.c_int, [vp, fp, fp, C.POINTER(i64)]),
        "goldq_replay_push": (C.c_int, [vp, i64, vp, vp, vp, vp, vp]),
        "goldq_replay_push_packed": (C.c_int, [vp, i64, vp]),
        "goldq_packed_bytes": (i64, [vp, i64]),
        "goldq_learn_prepare": (C.c_int, [vp]),
        "goldq_graph_builds": (i64, [vp]),
        "goldq_launch_floor": (C.c_int, [vp, i32, fp]),
        "goldq_time_forward": (C.c_int, [vp, i32, fp]),
        "goldq_resize_batch": (C.c_int, [vp, i32]),
        "goldq_batch_size": (i32, [vp]),
        "goldq_replay_len": (i64, [vp]),
        "goldq_replay_push_wait": (C.c_int, [vp]),
        "goldq_replay_stage_packed": (C.c_int, [vp, C.c_int64, vp]),
        "goldq_replay_commit": (C.c_int, [vp]),
        "goldq_host_alloc": (vp, [i64]),
        "goldq_host_free": (None, [vp]),
        "goldq_learn": (C.c_int, [vp, vp, u64]),
        "goldq_learn_device_indices": (C.c_int, [vp, vp]),
        "goldq_learn_many": (C.c_int, [vp, i32, u64]),
        "goldq_sampler_indices": (C.c_int, [i64, i32, u64, i32, ip]),
        "goldq_sync_target": (C.c_int, [vp]),
        "goldq_last_loss": (C.c_int, [vp, fp]),
        "goldq_predict": (C.c_int, [vp, C.c_int, C.c_int, i64, fp, fp]),
        "goldq_greedy_action": (C.c_int, [vp, C.c_int, i64, fp, ip]),
        "goldq_epsilon_greedy_action": (C.c_int, [vp, C.c_int, i64, fp, C.c_float, u64, ip]),
        "goldq_explore_plan": (C.c_int, [i64, i32, C.c_float, u64, up, ip]),
        "goldq_fit_batch": (C.c_int, [vp, i64, fp, fp]),
        "goldq_dp_unique_id": (C.c_int, [up]),
        "goldq_dp_init": (C.c_int, [vp, up, C.c_int, C.c_int]),
        "goldq_dp_mode": (C.c_int, [vp]),
        "goldq_sync": (C.c_int, [vp]),
        "goldq_stream": (vp, [vp]),
        "goldq_launch_count": (i64, [vp]),
        "goldq_debug_enable": (C.c_int, [vp, C.c_int]),
        "goldq_debug_fetch": (C.c_int, [vp, C.c_int, vp, i64]),
        "goldq_timer_start": (C.c_int, [vp]),
        "goldq_timer_stop": (C.c_int, [vp, fp]),
        "goldq_selftest_umma": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, fp]),
        "goldq_step_breakdown": (C.c_int, [vp, u64, i32, fp, ip, C.c_char_p, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.goldq_abi_version() != ABI_VERSION:
        raise ImportError("libgoldq.so ABI version mismatch: rebuild")
    _lib = L
    return L


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def default_config(**kw):
    cfg = GoldqConfig()
    lib().goldq_default_config(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise TypeError(f"unknown config field {k}")
        setattr(cfg, k, v)
    return cfg


def explore_plan(n, n_actions, epsilon, seed):
    """Host mirror of the device's epsilon-greedy draws: (explore[n] bool, random_action[n] int32)."""
    ex = np.empty(n, np.uint8)
    ra = np.empty(n, np.int32)
    rc = lib().goldq_explore_plan(n, n_actions, epsilon, seed, ex.ctypes.data_as(C.POINTER(C.c_uint8)),
                                  ra.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc:
        raise GoldqError(rc, "explore_plan: bad argument")
    return ex.astype(bool), ra


def sampler_indices(length, batch, seed, replica=0):
    out = np.empty(batch, np.int32)
    rc = lib().goldq_sampler_indices(length, batch, seed, replica, out.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc:
        raise GoldqError(rc, "sampler_indices: bad argument")
    return out


def selftest_umma(mode, A_bits, B_bits, M, N, K, nsplit=1):
    """A_bits/B_bits: uint16 bf16 bit patterns in the layout `mode` expects."""
    A_bits = np.ascontiguousarray(A_bits, np.uint16)
    B_bits = np.ascontiguousarray(B_bits, np.uint16)
    out = np.empty((M, N), np.float32)
    rc = lib().goldq_selftest_umma(mode, M, N, K, nsplit, A_bits.ctypes.data, B_bits.ctypes.data, _f(out))
    if rc:
        raise GoldqError(rc, lib().goldq_last_error(None).decode())
    return out


class Handle:
    """Thin owner of a goldq_t*; one per agent (or replica group) per GPU."""

    def __init__(self, cfg=None, **kw):
        self._L = lib()
        self.cfg = cfg if cfg is not None else default_config(**kw)
        self._h = self._L.goldq_create(C.byref(self.cfg))
        if not self._h:
            msg = self._L.goldq_last_error(None).decode()
            raise GoldqError(EINVAL, msg)
        self.P = int(self._L.goldq_param_count(self._h))
        self.R = int(self.cfg.n_replicas)
        self.B = int(self.cfg.batch_size)
        self.D = int(self.cfg.obs_dim)
        self.A = int(self.cfg.n_actions)

    def close(self):
        if getattr(self, "_h", None):
            self._L.goldq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise GoldqError(rc, self._L.goldq_last_error(self._h).decode())

    @property
    def path(self):
        return int(self._L.goldq_selected_path(self._h))

    # parameters / solver state -------------------------------------------
    def set_params(self, net, theta):
        theta = np.ascontiguousarray(theta, np.float32)
        assert theta.size == self.R * self.P, (theta.size, self.R, self.P)
        self._ck(self._L.goldq_set_params(self._h, net, _f(theta)))

    def get_params(self, net):
        out = np.empty(self.R * self.P, np.float32)
        self._ck(self._L.goldq_get_params(self._h, net, _f(out)))
        return out

    def set_adam_state(self, m, v, it):
        m = np.ascontiguousarray(m, np.float32)
        v = np.ascontiguousarray(v, np.float32)
        assert m.size == self.R * self.P and v.size == self.R * self.P
        self._ck(self._L.goldq_set_adam_state(self._h, _f(m), _f(v), it))

    def get_adam_state(self):
        m = np.empty(self.R * self.P, np.float32)
        v = np.empty(self.R * self.P, np.float32)
        it = C.c_int64(0)
        self._ck(self._L.goldq_get_adam_state(self._h, _f(m), _f(v), C.byref(it)))
        return m, v, it.value

    # replay ---------------------------------------------------------------
    def replay_push(self, s, a, r, s2, done):
        s = np.ascontiguousarray(s, np.float32)
        s2 = np.ascontiguousarray(s2, np.float32)
        a = np.ascontiguousarray(a, np.int32)
        r = np.ascontiguousarray(r, np.float32)
        done = np.ascontiguousarray(done, np.uint8)
        n = a.size // self.R
        assert s.size == self.R * n * self.D and s2.size == s.size and r.size == a.size == done.size
        self._ck(self._L.goldq_replay_push(self._h, n, s.ctypes.data, a.ctypes.data, r.ctypes.data,
                                           s2.ctypes.data, done.ctypes.data))

    def replay_push_raw(self, n, s_ptr, a_ptr, r_ptr, s2_ptr, done_ptr):
        self._ck(self._L.goldq_replay_push(self._h, n, s_ptr, a_ptr, r_ptr, s2_ptr, done_ptr))

    def packed_bytes(self, n):
        return int(self._L.goldq_packed_bytes(self._h, n))

    def replay_push_packed(self, n, packed_ptr):
        """One contiguous host buffer s|s2|a|r|done (goldq.h): ONE H2D copy."""
        self._ck(self._L.goldq_replay_push_packed(self._h, n, packed_ptr))

    def replay_stage_packed(self, n, packed_ptr):
        """Start the H2D copy of a packed batch; the replay memory changes at replay_commit()."""
        self._ck(self._L.goldq_replay_stage_packed(self._h, n, packed_ptr))

    def replay_commit(self):
        """Append the oldest staged batch to the replay memory (in stream order on the learn stream)."""
        self._ck(self._L.goldq_replay_commit(self._h))

    def pack(self, s, a, r, s2, done, out=None):
        """Pack the five arrays of n transitions into the layout goldq_replay_push_packed takes."""
        s = np.ascontiguousarray(s, np.float32).ravel(); s2 = np.ascontiguousarray(s2, np.float32).ravel()
        a = np.ascontiguousarray(a, np.int32).ravel(); r = np.ascontiguousarray(r, np.float32).ravel()
        done = np.ascontiguousarray(done, np.uint8).ravel()
        parts = [s.view(np.uint8), s2.view(np.uint8), a.view(np.uint8), r.view(np.uint8), done]
        total = sum(p.size for p in parts)
        if out is None:
            out = np.empty(total, np.uint8)
        assert out.size == total
        o = 0
        for p in parts:
            out[o:o + p.size] = p
            o += p.size
        return out

    def resize_batch(self, n):
        """Sequential.ResizeBatch: new batch size, same parameters / solver state / replay."""
        self._ck(self._L.goldq_resize_batch(self._h, int(n)))
        self.B = int(self._L.goldq_batch_size(self._h))
        self.cfg.batch_size = self.B

    def learn_prepare(self):
        self._ck(self._L.goldq_learn_prepare(self._h))

    def graph_builds(self):
        return int(self._L.goldq_graph_builds(self._h))

    def replay_push_wait(self):
        self._ck(self._L.goldq_replay_push_wait(self._h))

    def replay_len(self):
        return int(self._L.goldq_replay_len(self._h))

    # learn ------------------------------------------------------------------
    def learn(self, idx=None, seed=0):
        if idx is None:
            self._ck(self._L.goldq_learn(self._h, None, seed))
        else:
            idx = np.ascontiguousarray(idx, np.int32)
            assert idx.size == self.R * self.B
            self._ck(self._L.goldq_learn(self._h, idx.ctypes.data, seed))

    def learn_raw(self, idx_host_ptr, seed=0):
        self._ck(self._L.goldq_learn(self._h, idx_host_ptr, seed))

    def learn_many(self, n, seed=0):
        self._ck(self._L.goldq_learn_many(self._h, n, seed))

    def learn_device_indices(self, dev_ptr):
        self._ck(self._L.goldq_learn_device_indices(self._h, dev_ptr))

    def sync_target(self):
        self._ck(self._L.goldq_sync_target(self._h))

    def last_loss(self):
        out = np.empty(self.R, np.float32)
        self._ck(self._L.goldq_last_loss(self._h, _f(out)))
        return out

    # model surface ----------------------------------------------------------
    def predict(self, net, x, replica=0):
        x = np.ascontiguousarray(x, np.float32).reshape(-1, self.D)
        q = np.empty((x.shape[0], self.A), np.float32)
        self._ck(self._L.goldq_predict(self._h, net, replica, x.shape[0], _f(x), _f(q)))
        return q

    def greedy_action(self, x, replica=0):
        x = np.ascontiguousarray(x, np.float32).reshape(-1, self.D)
        a = np.empty(x.shape[0], np.int32)
        self._ck(self._L.goldq_greedy_action(self._h, replica, x.shape[0], _f(x),
                                             a.ctypes.data_as(C.POINTER(C.c_int32))))
        return a

    def epsilon_greedy_action(self, x, epsilon, seed, replica=0):
        """Agent.Action for a batch of states (vectorised environments)."""
        x = np.ascontiguousarray(x, np.float32).reshape(-1, self.D)
        a = np.empty(x.shape[0], np.int32)
        self._ck(self._L.goldq_epsilon_greedy_action(self._h, replica, x.shape[0], _f(x), epsilon, seed,
                                                     a.ctypes.data_as(C.POINTER(C.c_int32))))
        return a

    def fit_batch(self, x, y):
        x = np.ascontiguousarray(x, np.float32).reshape(-1, self.D)
        y = np.ascontiguousarray(y, np.float32).reshape(-1, self.A)
        if x.shape[0] != y.shape[0]:
            raise GoldqError(EINVAL, "fit_batch: x and y row counts differ")
        self._ck(self._L.goldq_fit_batch(self._h, x.shape[0], _f(x), _f(y)))

    # data parallel ----------------------------------------------------------
    @staticmethod
    def dp_unique_id():
        buf = (C.c_uint8 * 128)()
        rc = lib().goldq_dp_unique_id(buf)
        if rc:
            raise GoldqError(rc, lib().goldq_last_error(None).decode())
        return bytes(buf)

    def dp_init(self, uid, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._ck(self._L.goldq_dp_init(self._h, buf, rank, world))

    def dp_mode(self):
        """0: no group, 1: ncclAllReduce + update kernel, 2: peer-memory exchange inside the update kernel."""
        return int(self._L.goldq_dp_mode(self._h))

    def last_error(self):
        return self._L.goldq_last_error(self._h).decode()

    # plumbing ---------------------------------------------------------------
    def sync(self):
        self._ck(self._L.goldq_sync(self._h))

    def stream(self):
        return self._L.goldq_stream(self._h)

    def launch_count(self):
        return int(self._L.goldq_launch_count(self._h))

    def debug_enable(self, on=True):
        self._ck(self._L.goldq_debug_enable(self._h, 1 if on else 0))

    def debug_fetch(self, what):
        RB = self.R * self.B
        shape, dt = {
            DBG_Q_ONLINE: ((RB, self.A), np.float32), DBG_Q_TARGET: ((RB, self.A), np.float32),
            DBG_TD: ((RB,), np.float32), DBG_GRADS: ((self.R * self.P,), np.float32),
            DBG_INDICES: ((RB,), np.int32), DBG_BATCH_S: ((RB, self.D), np.float32),
            DBG_BATCH_A: ((RB,), np.int32), DBG_FF_STAMPS: ((self.B // 128 * 2, 96), np.int64),
            DBG_TRACE: ((4 + (1 << 17) * 12,), np.uint64),
        }[what]
        out = np.empty(shape, dt)
        self._ck(self._L.goldq_debug_fetch(self._h, what, out.ctypes.data, out.nbytes))
        return out

    def step_breakdown(self, seed=0):
        """[(kernel label, ms)] of one eager, serialised learn step."""
        ms = np.zeros(64, np.float32)
        n = C.c_int32(0)
        names = C.create_string_buffer(2048)
        self._ck(self._L.goldq_step_breakdown(self._h, seed, 64, _f(ms), C.byref(n), names, 2048))
        labels = names.value.decode().split("\n")
        return [(labels[i] if i < len(labels) else f"kernel{i}", float(ms[i])) for i in range(n.value)]

    def launch_floor_us(self, reps=50):
        us = C.c_float(0)
        self._ck(self._L.goldq_launch_floor(self._h, reps, C.byref(us)))
        return us.value

    def time_forward_us(self, reps=20):
        us = C.c_float(0)
        self._ck(self._L.goldq_time_forward(self._h, reps, C.byref(us)))
        return us.value

    def timer_start(self):
        self._ck(self._L.goldq_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float(0)
        self._ck(self._L.goldq_timer_stop(self._h, C.byref(ms)))
        return ms.value
