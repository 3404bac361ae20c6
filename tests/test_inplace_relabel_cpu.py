"""Host-side planning of the in-place relabelling (q1tsim_b200/csrc/planner.cpp, plan_inplace_relabel):
the Swap gates of the reference (swap.rs:78-88) are zero-byte relabels in the engine, and canonical index order is
restored before anything observes the state.  When no second column buffer fits into device memory the restore is
a sequence of TILE-CLOSED passes that may run with source == destination.  Checked here, without a GPU: every pass
only moves bits inside its tile, the tile holds the low coalescing bits, and the passes compose to the requested
permutation."""
import random

import pytest

from q1tsim_b200 import engine as E


def _check(dstpos, tile_bits=12, coalesce=3):
    n = len(dstpos)
    passes = E.plan_inplace_relabel(dstpos, tile_bits, coalesce)
    cur = list(range(n))                 # cur[p] = where the data that started at bit p is now
    for tile, dp in passes:
        assert len(tile) == min(tile_bits, n) and tile == sorted(set(tile))
        assert all(b in tile for b in range(min(coalesce, n)))         # 128-byte accesses
        assert sorted(dp) == list(range(n))
        for p in range(n):
            assert (dp[p] in tile) if p in tile else dp[p] == p        # tile-closed: a CTA writes what it has read
        assert dp != list(range(n))                                    # no empty passes
        cur = [dp[x] for x in cur]
    assert cur == list(dstpos)
    return len(passes)


@pytest.mark.parametrize("n,expected", [(5, 1), (12, 1), (13, 1), (20, 2), (30, 4), (33, 4), (36, 4), (40, 5)])
def test_qft_bit_reversal(n, expected):
    # QFT-n ends with Swap(i, n-1-i) (README.md:59-68 pattern): the relabel to undo is the bit reversal
    assert _check([n - 1 - p for p in range(n)]) == expected


def test_identity_needs_no_pass():
    assert _check(list(range(20))) == 0


def test_long_cycles_and_random_permutations():
    assert _check([(p + 1) % 40 for p in range(40)]) <= 6
    r = random.Random(1)
    for _ in range(400):
        n = r.randint(5, 40)
        dp = list(range(n))
        r.shuffle(dp)
        assert _check(dp, r.choice([8, 10, 12, 13]), r.choice([2, 3])) <= 12


def test_argument_checks():
    with pytest.raises(E.EngineError):
        E.plan_inplace_relabel([0, 0, 1, 2, 3])          # not a permutation
    with pytest.raises(E.EngineError):
        E.plan_inplace_relabel([1, 0, 2, 3], tile_bits=4)
