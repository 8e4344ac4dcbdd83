"""Diagnostic: L2 -> shared-memory streaming rate of one producer thread per CTA (selftest modes 40..43)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gold_b200 import _capi as G  # noqa: E402

z = np.zeros(16, np.uint16)
KINDS = {40: "TMA 2D boxes (1 KB pitch), same data in every CTA", 41: "TMA 2D boxes, own rows per CTA",
         42: "contiguous 16 KB bulk copies, same data", 43: "contiguous 16 KB bulk copies, own region"}
for ctas in (1, 32, 128, 148):
    for mode, name in KINDS.items():
        out = G.selftest_umma(mode, z, z, 128 * ctas, 2 * ctas if ctas > 1 else 2, 64)
        bpc = out.ravel()[0:2 * ctas:2]
        print(f"ctas={ctas:3d} {name:52s}: {np.median(bpc):6.1f} B/clk per SM (min {bpc.min():.1f}, max {bpc.max():.1f})", flush=True)
