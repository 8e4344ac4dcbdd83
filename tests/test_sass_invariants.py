"""Properties of the compiled kernels that the numbers in DESIGN.md depend on and that no numerical test can see
(read off `cuobjdump -sass` of the built library; no GPU needed).

* A kernel launched programmatically (griddepcontrol.wait = ACQBULK in SASS) must not load global memory before the
  wait: `__ldg` / `const __restrict__` loads are invariant loads for the compiler, which may hoist them above the wait,
  i.e. above the completion of the kernel that produces the data (round 2: wrong TD targets, a loss lagging one step).
* The pair kernels really issue cta_group::2 MMAs and TMA loads; the fused forward stays small enough for the
  instruction cache (the fully unrolled version, 6 k instructions, lost a third of its cycles to instruction fetch).
"""
import os
import re
import shutil
import subprocess

import pytest

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gold_b200", "lib", "libgoldq.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def kernels():
    if not os.path.exists(LIB) or not os.path.exists(CUOBJDUMP):
        pytest.skip("library or cuobjdump not available")
    out = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    ks, name = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            ks[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and name:
            ks[name].append(m.group(1).strip())
    assert ks, "no kernels found in the library"
    return ks


def test_no_global_load_precedes_the_dependency_wait(kernels):
    checked = 0
    for name, ins in kernels.items():
        waits = [i for i, x in enumerate(ins) if re.search(r"\bACQBULK\b", x)]
        if not waits:
            continue
        checked += 1
        early = [x for x in ins[:waits[0]] if re.search(r"\b(LDG|LD)\.", x) and "LDGSTS" not in x]
        if "wide_td_dz2_kernel" in name:
            # The one deliberate exception: TD/dZ2 reads what the forward it is launched behind does NOT produce -- the
            # batch's actions / rewards / done flags (3 loads; written by the gather, a full dependency of the forward)
            # and the W3^T rows those actions pick (8 loads; written by the previous update) -- ahead of the wait.  They
            # are ordinary cached loads (ld.global.ca), never the invariant kind the compiler may move.
            assert len(early) == 11, (name, early)
            assert all(".STRONG.SM" in x and "CONSTANT" not in x for x in early), (name, early)
            continue
        assert not early, f"{name}: global loads in front of griddepcontrol.wait: {early[:4]}"
    assert checked >= 6, checked      # forward(s), pair GEMMs, GEMMs, gather, TD, column sums, update


def test_pair_kernels_use_cta_group_2_and_stay_small(kernels):
    def find(sub):
        hits = [k for k in kernels if sub in k]
        assert hits, sub
        return hits
    for k in find("ff_forward2_kernel") + find("pair_gemm_kernel"):
        text = "\n".join(kernels[k])
        assert "UTCHMMA.2CTA" in text and "UTMALDG.2D.2CTA" in text, k
        assert "LDTM" in text and "UTMASTG" in text or "pair_gemm_kernelILi1" in k, k
    for k in find("ff_forward2_kernel"):
        assert len(kernels[k]) < 2600, (k, len(kernels[k]))      # rolled loops: ~1.8 k instructions (unrolled: 6 k)
    for k in find("umma_gemm_kernel"):
        assert "UTCHMMA" in "\n".join(kernels[k]), k


def test_epilogues_address_shared_memory_explicitly(kernels):
    """Staging tiles are written with STS / read with LDS, not generic ST.E / LD.E through a decayed pointer."""
    for k in [k for k in kernels if "ff_forward2_kernel" in k or "pair_gemm_kernelILi0" in k]:
        ins = kernels[k]
        generic_st = [x for x in ins if re.search(r"\bST\.E\.128\b", x)]
        sts = [x for x in ins if "STS.128" in x]
        assert len(sts) >= 16, (k, len(sts))
        # (row-major global stores of Q / masks / column sums remain: at most a handful of 16-byte generic stores)
        assert len(generic_st) <= 8, (k, generic_st[:4])
