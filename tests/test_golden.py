"""Committed fixtures (tests/golden/, made by tests/golden/make_golden.py).

CPU: the oracle must still reproduce them bit for bit (guards the checker itself), and the
literals of the reference's own tests must hold.  GPU: the CUDA paths, driven through the C ABI
from the fixture's inputs only, must land on the fixture's outputs."""
import json
import os

import numpy as np
import pytest

from util import OracleAgent, assert_params_close, rel_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["cartpole_b20", "odd_b16"]


def _load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name):
    g = _load(name)
    dims, B = tuple(int(x) for x in g["dims"]), int(g["batch"])
    ag = OracleAgent(dims, B, g["theta_init"], g["theta_t_init"])
    ag.push(g["s"], g["a"], g["r"], g["s2"], g["done"])
    for step in range(3):
        o, _ = ag.learn(g[f"idx{step}"], with64=False)
        for k in ("q_online", "q_target", "td", "grads"):
            np.testing.assert_array_equal(np.asarray(o[k], np.float32), g[f"{k}{step}"], err_msg=f"{k} step {step}")
        assert np.float32(o["loss"]) == g[f"loss{step}"]
        np.testing.assert_array_equal(ag.theta, g[f"theta{step}"])
        np.testing.assert_array_equal(ag.m, g[f"m{step}"])
        np.testing.assert_array_equal(ag.v, g[f"v{step}"])
        if step == 1:
            ag.sync_target()


def test_reference_test_literals():
    from oracle import binding as O
    k = json.load(open(os.path.join(GOLD, "reference_kats.json")))
    assert O.amax(np.array(k["amax"]["input"], np.float32)) == k["amax"]["expect"]
    rows = [np.array(r, np.float32).reshape(1, -1) for r in k["concat"]["rows"]]
    c = np.concatenate(rows, 0)                     # the gather writes batch-major rows: same stacking
    assert list(c.shape) == k["concat"]["expect_shape"] and c.ravel().tolist() == k["concat"]["expect_flat"]
    from gold_b200 import synth
    idx = synth.sample_indices(k["sample_len"]["pushed"], k["sample_len"]["batch"], 7)
    assert len(idx) == k["sample_len"]["expect_len"] == len(set(idx.tolist()))


@pytest.mark.gpu
@pytest.mark.parametrize("name,path", [("cartpole_b20", "narrow"), ("cartpole_b20", "generic"), ("odd_b16", "generic")])
def test_cuda_path_matches_golden(name, path):
    from gold_b200 import _capi as G
    g = _load(name)
    dims, B = tuple(int(x) for x in g["dims"]), int(g["batch"])
    kw = dict(path=G.PATH_GENERIC) if path == "generic" else {}
    h = G.Handle(obs_dim=dims[0], hidden1=dims[1], hidden2=dims[2], n_actions=dims[3], batch_size=B,
                 replay_capacity=len(g["a"]), **kw)
    assert h.path == (G.PATH_GENERIC if path == "generic" else G.PATH_NARROW)
    h.debug_enable(True)
    h.set_params(G.NET_ONLINE, g["theta_init"]); h.set_params(G.NET_TARGET, g["theta_t_init"])
    h.replay_push(g["s"], g["a"], g["r"], g["s2"], g["done"])
    for step in range(3):
        h.learn(g[f"idx{step}"])
        # forward arithmetic is the reference's: bit-exact
        np.testing.assert_array_equal(h.debug_fetch(G.DBG_Q_ONLINE).reshape(B, -1), g[f"q_online{step}"])
        np.testing.assert_array_equal(h.debug_fetch(G.DBG_Q_TARGET).reshape(B, -1), g[f"q_target{step}"])
        np.testing.assert_array_equal(h.debug_fetch(G.DBG_TD).reshape(B), g[f"td{step}"])
        assert abs(float(h.last_loss()[0]) - float(g[f"loss{step}"])) <= 1e-5 * abs(float(g[f"loss{step}"]))
        gr = h.debug_fetch(G.DBG_GRADS)[: len(g[f"grads{step}"])]
        assert rel_err(gr, g[f"grads{step}"]) <= 1e-5
        assert_params_close(h.get_params(G.NET_ONLINE), g[f"theta{step}"], g[f"grads{step}"])
        # keep the trajectories from drifting apart through ill-conditioned elements
        h.set_params(G.NET_ONLINE, g[f"theta{step}"]); h.set_adam_state(g[f"m{step}"], g[f"v{step}"], step + 1)
        if step == 1:
            h.sync_target()
    h.close()
