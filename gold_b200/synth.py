"""Synthetic inputs for tests and bench.

Everything comes from numpy PCG64 streams with fixed seeds so that the oracle
and the CUDA path see identical weights, replay contents and sampled indices.
"""
import numpy as np

CARTPOLE = (4, 24, 24, 2)
WIDE = (128, 512, 512, 18)

SEED_DATA = 1234
SEED_ONLINE = 1
SEED_TARGET = 2
SEED_INDEX0 = 100


def nparams(dims):
    D, H1, H2, A = dims
    return D * H1 + H1 + H1 * H2 + H2 + H2 * A + A


def offsets(dims):
    """Flat learnable layout [W1,b1,W2,b2,W3,b3]; W is (in x out) row-major
    (reference: vendor/github.com/aunum/goro/pkg/v1/layer/fc.go:98-101,164-169)."""
    D, H1, H2, A = dims
    o = {}
    p = 0
    for name, n in (("w1", D * H1), ("b1", H1), ("w2", H1 * H2), ("b2", H2),
                    ("w3", H2 * A), ("b3", A)):
        o[name] = (p, p + n)
        p += n
    return o


def glorot_params(dims, seed):
    """GlorotN(1): N(0, 2/(n1+n2)) for W (in x out) and N(0, 2/(1+out)) for the
    (1 x out) bias (reference: vendor/gorgonia.org/gorgonia/weights.go:266-293,
    goro layer/fc.go:76-81)."""
    D, H1, H2, A = dims
    rng = np.random.Generator(np.random.PCG64(seed))
    parts = []
    for fin, fout in ((D, H1), (H1, H2), (H2, A)):
        parts.append(rng.normal(0.0, np.sqrt(2.0 / (fin + fout)), fin * fout))
        parts.append(rng.normal(0.0, np.sqrt(2.0 / (1 + fout)), fout))
    return np.concatenate(parts).astype(np.float32)


def replay(dims, n, seed=SEED_DATA, p_done=0.05):
    """n transitions in ARRIVAL order (oldest first)."""
    D, _, _, A = dims
    rng = np.random.Generator(np.random.PCG64(seed))
    s = rng.uniform(-1.0, 1.0, (n, D)).astype(np.float32)
    s2 = rng.uniform(-1.0, 1.0, (n, D)).astype(np.float32)
    a = rng.integers(0, A, n).astype(np.int32)
    r = rng.uniform(0.0, 1.0, n).astype(np.float32)
    done = (rng.uniform(0.0, 1.0, n) < p_done).astype(np.uint8)
    return s, a, r, s2, done


def sample_indices(n, batch, step):
    """First `batch` entries of a permutation of n: distinct, uniform — what
    Memory.Sample draws (reference: pkg/v1/agent/deepq/memory.go:54-69).  Index i
    means the i-th newest event."""
    rng = np.random.Generator(np.random.PCG64(SEED_INDEX0 + step))
    if batch * 4 < n:
        # same distribution, O(batch) instead of O(n)
        return rng.choice(n, size=batch, replace=False).astype(np.int32)
    return rng.permutation(n)[:batch].astype(np.int32)


def atari_params(C, H, W, A, seed):
    """Initial parameters of gold's Atari conv policy in learnable order [f1, f2, f3, W4, b4, W5, b5]: the filters
    GlorotU(1) (goro layer/conv2d.go:104-106: uniform on +-sqrt(6 / (fan_in + fan_out)) with gorgonia's fans for a
    4-d shape, weights.go:266-293), the FC layers GlorotN(1) as in glorot_params."""
    rng = np.random.Generator(np.random.PCG64(seed))
    h1, w1 = (H + 2 - 8) // 4 + 1, (W + 2 - 8) // 4 + 1
    h2, w2 = (h1 + 2 - 4) // 2 + 1, (w1 + 2 - 4) // 2 + 1
    F = 64 * h2 * w2
    out = []
    for o, i, k in ((32, C, 8), (64, 32, 4), (64, 64, 3)):
        lim = np.sqrt(6.0 / (i * k * k + o * k * k))
        out.append(rng.uniform(-lim, lim, o * i * k * k))
    for fin, fout in ((F, 512), (512, A)):
        out.append(rng.standard_normal(fin * fout) * np.sqrt(2.0 / (fin + fout)))
        out.append(rng.standard_normal(fout) * np.sqrt(2.0 / (1 + fout)))
    return np.concatenate(out).astype(np.float32)
