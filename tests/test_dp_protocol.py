"""Model check of the data-parallel exchange protocol (gold_b200/csrc/wide.cuh, PeerSet; kernel:
wide_update_kernel<true, true> / dp_exchange4 in kernels_wide.cu).

Every gradient element travels as a {value, epoch} word.  Per element, the same thread on every
rank: push -> (owner only) reduce + broadcast -> apply; a rank starts step e+1 only when all its
threads have finished step e (stream order of the kernels).  The claims the CUDA code relies on,
checked here over random interleavings:
  safety    a word is never overwritten before its reader has consumed it, and a reader never
            consumes a word of the wrong epoch (so no flags / fences are needed beside the tag);
  liveness  some action is always enabled until every rank has finished every step;
  result    every rank applies the rank-ordered sum of all ranks' values;
and that the check has teeth: drop the "finish step e before starting e+1" rule and it breaks.
(The CUDA code alternates two windows by epoch parity; the model shows one would already be safe,
so the second is margin, not a requirement.)"""
import random

import pytest


class Word:
    __slots__ = ("val", "epoch", "consumed")

    def __init__(self):
        self.val, self.epoch, self.consumed = 0.0, 0, True      # zeroed windows; epochs start at 1


class Overwrite(Exception):
    pass


def simulate(world, groups, epochs, seed, parities=2, kernel_order=True):
    rnd = random.Random(seed)
    owner = lambda g: g * world // groups                            # contiguous ranges, like chunk_f
    grad = lambda r, g, e: float((r + 1) * 1000 + g * 10 + e)        # exact in fp32/fp64
    rs = [[[[Word() for _ in range(groups)] for _ in range(world)] for _ in range(parities)] for _ in range(world)]
    ag = [[[Word() for _ in range(groups)] for _ in range(parities)] for _ in range(world)]
    epoch = [1] * world
    phase = [[0] * groups for _ in range(world)]                     # 0 idle, 1 pushed, 2 applied
    reduced = [[False] * groups for _ in range(world)]
    applied = {}

    def write(w, val, e):
        if not w.consumed:
            raise Overwrite(f"word of epoch {w.epoch} overwritten by epoch {e} before it was read")
        w.val, w.epoch, w.consumed = val, e, False

    def enabled():
        acts = []
        for r in range(world):
            if epoch[r] > epochs:
                continue
            e, p = epoch[r], epoch[r] % parities
            if all(ph == 2 for ph in phase[r]):
                acts.append(("advance", r, -1))
                continue
            if not kernel_order and all(ph >= 1 for ph in phase[r]):
                acts.append(("advance", r, -1))           # (broken variant: next step before this one is applied)
            for g in range(groups):
                if phase[r][g] == 0:
                    acts.append(("push", r, g))
                elif phase[r][g] == 1:
                    if owner(g) == r and not reduced[r][g] and all(rs[r][p][s][g].epoch == e for s in range(world)):
                        acts.append(("reduce", r, g))
                    if ag[r][p][g].epoch == e and (owner(g) != r or reduced[r][g]):
                        acts.append(("apply", r, g))
        return acts

    steps = 0
    while any(ep <= epochs for ep in epoch):
        acts = enabled()
        assert acts, f"deadlock at epochs {epoch}"
        kind, r, g = rnd.choice(acts)
        e, p = epoch[r], epoch[r] % parities
        if kind == "push":
            write(rs[owner(g)][p][r][g], grad(r, g, e), e)
            phase[r][g] = 1
        elif kind == "reduce":
            acc = 0.0
            for s in range(world):                                    # rank order
                w = rs[r][p][s][g]
                assert w.epoch == e
                acc += w.val
                w.consumed = True
            for k in range(world):
                write(ag[k][p][g], acc, e)
            reduced[r][g] = True
        elif kind == "apply":
            w = ag[r][p][g]
            assert w.epoch == e
            w.consumed = True
            applied[(r, g, e)] = w.val
            phase[r][g] = 2
        else:
            epoch[r] += 1
            phase[r] = [0] * groups
            reduced[r] = [False] * groups
        steps += 1
        assert not kernel_order or max(epoch) - min(epoch) <= 1, "a rank ran more than one step ahead of a peer"
    for (r, g, e), v in applied.items():
        assert v == sum(grad(s, g, e) for s in range(world))
    assert len(applied) == world * groups * epochs
    return steps


@pytest.mark.parametrize("world,groups", [(2, 4), (3, 5), (4, 8), (8, 8)])
def test_exchange_protocol_is_safe_and_live(world, groups):
    for seed in range(40):
        simulate(world, groups, epochs=6, seed=seed)


@pytest.mark.parametrize("world,groups", [(2, 4), (4, 8)])
def test_one_window_is_already_safe(world, groups):
    for seed in range(40):
        simulate(world, groups, epochs=6, seed=seed, parities=1)


def test_kernel_order_is_what_makes_it_safe():
    """Without 'all threads of step e finish before step e+1 starts' a fast rank overwrites words its
    peers have not consumed: the model must catch that, or the checks above prove nothing."""
    broke = 0
    for seed in range(100):
        try:
            simulate(3, 4, epochs=4, seed=seed, parities=1, kernel_order=False)
        except (Overwrite, AssertionError, KeyError):
            broke += 1
    assert broke > 50
