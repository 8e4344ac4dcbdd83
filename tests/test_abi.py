"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU
and exports exactly the symbols include/goldq.h declares; pure-host entry points
work; compute entry points fail loudly without a device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from gold_b200 import _capi as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "goldq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(goldq_[a-z0-9_]+)\s*\(", src)))


def test_header_binding_and_library_agree():
    hs = header_symbols()
    assert hs == sorted(G.SYMBOLS), set(hs) ^ set(G.SYMBOLS)
    L = C.CDLL(G.LIB_PATH)
    for s in hs:
        assert hasattr(L, s), f"libgoldq.so does not export {s}"


def test_abi_version_and_default_config():
    assert G.lib().goldq_abi_version() == G.ABI_VERSION
    c = G.default_config()
    # gold's defaults: agent.go:60-65, policy.go:32-38
    assert (c.obs_dim, c.hidden1, c.hidden2, c.n_actions, c.batch_size) == (4, 24, 24, 2, 20)
    assert abs(c.gamma - 0.95) < 1e-7 and c.lr == 0.0005 and c.beta1 == 0.9 and c.beta2 == 0.999 and c.eps == 1e-8
    assert c.replay_capacity == 10_000_000 and c.n_replicas == 1


def test_config_struct_layout_matches_header():
    # field order in the ctypes mirror == field order in the header
    src = open(os.path.join(ROOT, "include", "goldq.h")).read()
    body = src[src.index("typedef struct goldq_config {"):src.index("} goldq_config_t;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.findall(r"\*?\s*([a-z0-9_]+)$", part.strip())[0])
    assert names == [f[0] for f in G.GoldqConfig._fields_]


def test_host_sampler_is_a_keyed_permutation():
    for n in (1, 2, 7, 100, 4097):
        full = G.sampler_indices(n, n, 99)
        assert sorted(full.tolist()) == list(range(n))
    a = G.sampler_indices(1 << 20, 4096, 1)
    b = G.sampler_indices(1 << 20, 4096, 2)
    assert len(set(a.tolist())) == 4096 and (a != b).mean() > 0.99
    # roughly uniform: mean close to n/2
    assert abs(a.astype(np.float64).mean() / (1 << 20) - 0.5) < 0.03
    # replicas draw different streams
    assert (G.sampler_indices(1000, 20, 5, 0) != G.sampler_indices(1000, 20, 5, 1)).any()


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(G.GoldqError):
        G.Handle()


def test_bad_config_rejected_before_touching_the_device():
    with pytest.raises(G.GoldqError):
        G.Handle(batch_size=0)
    with pytest.raises(G.GoldqError):
        G.Handle(abi_version=99)


def test_explore_plan_is_uniform_and_deterministic():
    """Host mirror of the batched epsilon-greedy draws (goldq_explore_plan; no GPU needed)."""
    from gold_b200 import _capi as G
    ex, ra = G.explore_plan(200000, 6, 0.25, 12345)
    ex2, ra2 = G.explore_plan(200000, 6, 0.25, 12345)
    assert (ex == ex2).all() and (ra == ra2).all()
    assert abs(ex.mean() - 0.25) < 5e-3                       # ~5 sigma of a Bernoulli(0.25) over 2e5 draws
    counts = np.bincount(ra, minlength=6) / len(ra)
    assert ra.min() >= 0 and ra.max() == 5 and np.abs(counts - 1 / 6).max() < 5e-3
    assert abs(np.corrcoef(ex.astype(float), ra.astype(float))[0, 1]) < 0.01
    assert not G.explore_plan(1000, 3, 0.0, 7)[0].any() and G.explore_plan(1000, 3, 1.0, 7)[0].all()
    assert (G.explore_plan(1000, 3, 0.5, 7)[0] != G.explore_plan(1000, 3, 0.5, 8)[0]).any()
