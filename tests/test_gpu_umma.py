"""Stand-alone checks of the hand-written tcgen05 GEMM tiles (umma_gemm_kernel)
against numpy on the same bf16 operands: all operand layouts (KK, MM, MK), the 32-wide N tile,
ragged shapes (TMA zero-fill + store guards) and split-K slices."""
import numpy as np
import pytest

from gold_b200 import _capi as G

pytestmark = pytest.mark.gpu


def bf16_bits(x):
    """float32 -> bf16 (round to nearest even) bit patterns, and the rounded values."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    back = (r.astype(np.uint32) << 16).view(np.float32)
    return r, back


@pytest.mark.parametrize("mode,M,N,K,nsplit", [
    (0, 128, 128, 64, 1), (0, 128, 128, 128, 1), (0, 256, 512, 512, 1), (0, 200, 136, 72, 1),
    (0, 8192, 512, 128, 1), (0, 256, 256, 1024, 4),
    (2, 128, 32, 64, 1), (2, 384, 32, 512, 1), (2, 100, 18, 512, 1),
    (1, 128, 128, 64, 1), (1, 128, 128, 256, 1), (1, 512, 512, 1024, 4), (1, 128, 512, 8192, 32),
    (1, 136, 200, 100, 2),
    (3, 128, 32, 64, 1), (3, 512, 32, 8192, 32), (3, 256, 32, 1000, 3),
    (4, 128, 128, 128, 1), (4, 8192, 512, 512, 1), (4, 200, 136, 72, 1), (5, 256, 64, 512, 1), (5, 100, 64, 128, 1),
    # B tile multicast over clusters of 4 CTAs along M (M tiles must divide by 4)
    (10, 512, 128, 64, 1), (10, 8192, 512, 512, 1), (10, 1024, 256, 1024, 2), (11, 512, 512, 1024, 4),
    (11, 512, 128, 192, 1), (14, 512, 128, 128, 1), (14, 8192, 512, 512, 1), (14, 1024, 136, 72, 1),
    # 128x256 tiles, two stages
    (20, 256, 256, 128, 1), (20, 8192, 512, 512, 1), (20, 200, 264, 72, 1), (24, 256, 256, 128, 1), (24, 8192, 512, 512, 1),
])
def test_umma_gemm_matches_numpy(mode, M, N, K, nsplit):
    api_mode, mode = mode, mode % 10
    rng = np.random.Generator(np.random.PCG64(M * 7 + N * 3 + K + mode))
    a = rng.normal(0, 1, (M, K)).astype(np.float32)
    b = rng.normal(0, 1, (N, K)).astype(np.float32)
    abits, ar = bf16_bits(a)
    bbits, br = bf16_bits(b)
    ref = ar.astype(np.float64) @ br.astype(np.float64).T
    if mode == 1:
        if (M % 8) or (N % 8):
            pytest.skip("MN-major operands need a 16-byte row pitch")
        abits = np.ascontiguousarray(abits.T)      # [K][M]
        bbits = np.ascontiguousarray(bbits.T)      # [K][N]
    elif mode == 3:
        abits = np.ascontiguousarray(abits.T)      # A MN-major [K][M], B stays K-major [N][K]
    elif mode in (4, 5):
        if (K % 8) or (N % 8):
            pytest.skip("16-byte row pitch")
        bbits = np.ascontiguousarray(bbits.T)      # A K-major [M][K], B MN-major [K][N]
    elif K % 8:
        pytest.skip("K-major operands need a 16-byte row pitch")
    if mode == 2 and N < 32:
        # the 32-wide tile reads 32 B rows: pad B with zero rows like W3^T is padded
        bpad = np.zeros((32, K), np.uint16); bpad[:N] = bbits; bbits = bpad
        ref_full = np.zeros((M, 32)); ref_full[:, :N] = ref; ref = ref_full; N = 32
    out = G.selftest_umma(api_mode, abits, bbits, M, N, K, nsplit)
    scale = np.sqrt(K)
    np.testing.assert_allclose(out, ref, rtol=2e-5, atol=2e-5 * scale)
