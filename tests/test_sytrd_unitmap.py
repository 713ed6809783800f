"""Index arithmetic of the tridiagonalisation kernel (rsoptsc_b200/csrc/sytrd_kernel.cuh, struct UnitMap),
restated in Python and checked exhaustively on small cases: the static assignment of strips to warps
and the slot numbering of the row partial sums must tile the work exactly once.

    unit u of block row b: first_of_row(b) = SPT b (b + 1) / 2, block row b has SPT (b + 1) strips
    owner(u)  = floor(u nw / total),  nw = min(warps, total)
    start(g)  = ceil(g total / nw)
    row-sum slot of (warp g, block row b) = g + b
"""
import itertools

import pytest

SPT = 4


def first_of_row(b):
    return SPT * b * (b + 1) // 2


def start(g, total, nw):
    return (g * total + nw - 1) // nw


def owner(u, total, nw):
    return u * nw // total


@pytest.mark.parametrize("mt,ndot,warps", list(itertools.product([1, 2, 3, 7, 20], [0, 5, 64], [3, 16, 2368])))
def test_runs_tile_the_units(mt, ndot, warps):
    strips = first_of_row(mt)
    total = strips + ndot
    nw = min(warps, total)
    seen = [0] * total
    for g in range(nw):
        s, e = start(g, total, nw), start(g + 1, total, nw)
        assert e > s                                   # every warp below nw owns at least one unit
        for u in range(s, e):
            assert owner(u, total, nw) == g
            seen[u] += 1
    assert seen == [1] * total
    assert start(nw, total, nw) == total


@pytest.mark.parametrize("mt,ndot,warps", list(itertools.product([1, 2, 5, 12, 31], [0, 17], [2, 16, 150, 2368])))
def test_row_sum_slots_are_unique_and_complete(mt, ndot, warps):
    strips = first_of_row(mt)
    total = strips + ndot
    nw = min(warps, total)
    written = {}
    for g in range(nw):
        for u in range(start(g, total, nw), min(start(g + 1, total, nw), strips)):
            b = 0
            while first_of_row(b + 1) <= u:
                b += 1
            written.setdefault(g + b, set()).add((g, b))
    # one (warp, block row) pair per slot
    assert all(len(v) == 1 for v in written.values())
    assert max(written) <= nw + mt
    # the reader of block row b visits exactly the slots that were written for it
    for b in range(mt):
        s0 = first_of_row(b)
        g0, g1 = owner(s0, total, nw), owner(s0 + SPT * (b + 1) - 1, total, nw)
        for g in range(g0, g1 + 1):
            assert written.get(g + b) == {(g, b)}
        others = [slot for slot, v in written.items() if next(iter(v))[1] == b and not (g0 <= next(iter(v))[0] <= g1)]
        assert not others
