"""Diagnostic: tcgen05.mma issue/completion rate (selftest modes 30/31); not a bench."""
import sys
import numpy as np
sys.path.insert(0, '.')
from gold_b200 import _capi as G
z = np.zeros(16, np.uint16)
for ctas in (1, 148):
    for mode, name in ((30, 'B MN-major'), (31, 'B K-major')):
        for N in (64, 128, 256):
            for naccs in (1, 2):
                out = G.selftest_umma(mode, z, z, 128 * ctas, N, 256, naccs)
                print(f"ctas={ctas:3d} {name} N={N:3d} accs={naccs}: issue {out.flat[0]:.1f} cyc/MMA, issue+drain {out.flat[1]:.1f} cyc/MMA", flush=True)
