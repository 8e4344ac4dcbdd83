"""Compares a dump of baseline/go_harness (real Gorgonia) with a tests/golden fixture (the oracle).

usage: python baseline/go_harness/compare.py tests/golden/cartpole_b20.npz /tmp/ref_dump
Expectation if the oracle restates the reference exactly: every array bit-identical (q_target
only on rows whose event is not terminal: the reference does not evaluate the target net there)."""
import os
import sys

import numpy as np


def main(npz, dump):
    g = np.load(npz)
    B = int(g["batch"]); N = len(g["a"]); worst = 0.0; ok = True
    for step in range(3):
        done = g["done"][N - 1 - g[f"idx{step}"]].astype(bool)
        for k in ("q_online", "q_target", "td", "loss", "grads", "theta"):
            ref = np.fromfile(os.path.join(dump, f"{k}{step}.f32"), np.float32)
            mine = np.asarray(g[f"{k}{step}"], np.float32).ravel()
            if k == "q_target":
                keep = np.repeat(~done, len(mine) // B)
                ref, mine = ref[keep], mine[keep]
            same = np.array_equal(ref, mine)
            err = float(np.abs(ref.astype(np.float64) - mine).max() / max(np.abs(ref).max(), 1e-30))
            worst = max(worst, err); ok &= same
            print(f"step {step} {k:9s} {'bit-identical' if same else f'DIFFERS, max rel err {err:.3e}'}")
    print("PINNED: oracle == executed reference" if ok else f"not bit-identical; worst relative error {worst:.3e}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1], sys.argv[2]))
