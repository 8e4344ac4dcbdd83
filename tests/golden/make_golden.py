"""Generates the fixtures in this directory (run from the repo root: python tests/golden/make_golden.py).

What they are: inputs and outputs of three consecutive Agent.Learn steps (target sync after the
second) computed by the fp32 CPU oracle (oracle/goldq_oracle.c) on seeded synthetic inputs, for
  cartpole_b20   the reference's own CPU-runnable configuration (4->24->24->2, batch 20), and
  odd_b16        a ragged shape (5->17->33->3, batch 16) that only the generic path accepts.
What they are NOT: output of the Go reference.  Gorgonia cannot run in this image (no Go
toolchain), so these vectors pin the oracle against drift and give the CUDA paths a committed
target; `reference_kats.json` holds the literals the reference's own tests assert for this path
(pkg/v1/dense/op_test.go:12-31, pkg/v1/agent/deepq/memory_test.go:28).  baseline/go_harness/
dumps the same cases from real Gorgonia for whoever has Go; DESIGN.md section 2.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {"cartpole_b20": ((4, 24, 24, 2), 20, 64), "odd_b16": ((5, 17, 33, 3), 16, 48)}


def make(name):
    from gold_b200 import synth
    from util import OracleAgent
    dims, B, N = CASES[name]
    theta, theta_t = synth.glorot_params(dims, 1), synth.glorot_params(dims, 2)
    s, a, r, s2, done = synth.replay(dims, N, seed=1234)
    ag = OracleAgent(dims, B, theta, theta_t)
    ag.push(s, a, r, s2, done)
    out = {"dims": np.array(dims, np.int32), "batch": np.int32(B), "theta_init": theta, "theta_t_init": theta_t,
           "s": s, "a": a, "r": r, "s2": s2, "done": done, "gamma": np.float32(0.95), "lr": np.float64(5e-4)}
    for step in range(3):
        idx = synth.sample_indices(N, B, 100 + step).astype(np.int32)
        o, _ = ag.learn(idx, with64=False)
        out[f"idx{step}"] = idx
        for k in ("q_online", "q_target", "td", "grads"):
            out[f"{k}{step}"] = np.asarray(o[k], np.float32)
        out[f"loss{step}"] = np.float32(o["loss"])
        out[f"theta{step}"] = ag.theta.copy()
        out[f"m{step}"] = ag.m.copy()
        out[f"v{step}"] = ag.v.copy()
        if step == 1:
            ag.sync_target()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def export_bin(name, root):
    """Raw little-endian inputs for baseline/go_harness (not committed: derived from the .npz)."""
    g = np.load(os.path.join(HERE, name + ".npz"))
    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    dims, B, N = CASES[name]
    np.array(list(dims) + [B, N, 3], "<i4").tofile(os.path.join(d, "meta.i32"))
    np.array([g["gamma"]], "<f4").tofile(os.path.join(d, "gamma.f32"))
    for k in ("s", "s2", "r", "theta_init", "theta_t_init"):
        np.asarray(g[k], "<f4").tofile(os.path.join(d, k + ".f32"))
    for k in ("a", "done", "idx0", "idx1", "idx2"):
        np.asarray(g[k]).astype("<i4").tofile(os.path.join(d, k + ".i32"))


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--bin":
        for n in CASES:
            export_bin(n, sys.argv[2])
        print("exported", sorted(CASES), "to", sys.argv[2])
        sys.exit(0)
    for n in CASES:
        make(n)
    kats = {"amax": {"input": [5, 8, 3, 2], "expect": 8, "ref": "pkg/v1/dense/op_test.go:12-18"},
            "concat": {"rows": [list(range(0, 5)), list(range(5, 10)), list(range(10, 15))], "expect_shape": [3, 5],
                       "expect_flat": list(range(15)), "ref": "pkg/v1/dense/op_test.go:20-31"},
            "sample_len": {"pushed": 10, "batch": 5, "expect_len": 5, "ref": "pkg/v1/agent/deepq/memory_test.go:15-35"}}
    json.dump(kats, open(os.path.join(HERE, "reference_kats.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(HERE)))
