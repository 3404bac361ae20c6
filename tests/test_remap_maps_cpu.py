"""The index maps of the qubit remap, restated in numpy: the in-place swap of `group_swap_kernel` (element (r, l) goes to
(r', l') with the k bit pairs exchanged, an involution) and the read-only form of it that `ladder_kernel<.., REMOTE>` and
`group_gather_kernel` use (kernels.h RemoteGather: element l of rank r is read from rank r[gb := l_lp] at index
l[lp := r_gb]).  Both must describe the same permutation of the P * 2^n amplitudes, for every choice of bit pairs."""
import itertools

import numpy as np
import pytest


def swap_remap(state, gb, lp):
    """state[r, l] -> new[r', l'] = state[r, l]: rank bit gb[j] and index bit lp[j] trade places"""
    P, N = state.shape
    r, l = np.meshgrid(np.arange(P), np.arange(N), indexing="ij")
    r2, l2 = r.copy(), l.copy()
    for g, p in zip(gb, lp):
        rb, lb = (r >> g) & 1, (l >> p) & 1
        r2 = (r2 & ~(1 << g)) | (lb << g)
        l2 = (l2 & ~(1 << p)) | (rb << p)
    out = np.empty_like(state)
    out[r2, l2] = state[r, l]
    return out


def gather_remap(state, gb, lp):
    """what every rank reads: new[r, l] = state[r with gb := l_lp, l with lp := r_gb]"""
    P, N = state.shape
    out = np.empty_like(state)
    lpmask = sum(1 << p for p in lp)
    for r in range(P):
        l = np.arange(N)
        src_rank = np.full(N, r & ~sum(1 << g for g in gb))
        mybits = 0
        for g, p in zip(gb, lp):
            src_rank |= ((l >> p) & 1) << g
            mybits |= ((r >> g) & 1) << p
        out[r] = state[src_rank, (l & ~lpmask) | mybits]
    return out


@pytest.mark.parametrize("g,n", [(1, 5), (2, 6), (3, 7)])
def test_gather_form_is_the_swap_involution(g, n):
    P, N = 1 << g, 1 << n
    state = np.arange(P * N).reshape(P, N)
    for k in range(1, g + 1):
        for gb in itertools.permutations(range(g), k):
            for lp in [tuple(range(n - k, n)), tuple(range(k)), tuple(reversed(range(1, k + 1))), (n - 1, 0, 2)[:k]]:
                a = swap_remap(state, gb, lp)
                b = gather_remap(state, gb, lp)
                assert np.array_equal(a, b), (gb, lp)
                assert np.array_equal(swap_remap(a, gb, lp), state)          # an involution
