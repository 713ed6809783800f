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


# ---- multi-GPU map (sytrd_mg.cuh): block row I (64 rows, absolute) belongs to rank (I // 2) % P for the whole
# factorisation; per column the rank's units are the strips of its block rows at or below Bmin --------------------
def mg_tables(mB, P, rank):
    own = [I for I in range(mB) if (I >> 1) % P == rank]
    pref = [0]
    for I in own:
        pref.append(pref[-1] + SPT * (I + 1))
    lfirst, l = [], 0
    for I in range(mB + 1):
        while l < len(own) and own[l] < I:
            l += 1
        lfirst.append(l)
    return own, pref, lfirst


def mg_F(pref, l0, Bmin, l):
    return pref[l] - pref[l0] - SPT * Bmin * (l - l0)


def mg_local_row(own, pref, l0, nl, Bmin, P, u):
    import math
    bf = own[l0] - Bmin
    b1 = bf + 1.0
    l = l0 + int((math.sqrt(b1 * b1 + 2.0 * P * u / SPT) - b1) / P)
    l = max(l0, min(l, nl - 1))
    steps = 0
    while l > l0 and mg_F(pref, l0, Bmin, l) > u:
        l -= 1
        steps += 1
    while l + 1 < nl and mg_F(pref, l0, Bmin, l + 1) <= u:
        l += 1
        steps += 1
    return l, steps


@pytest.mark.parametrize("mB,P", list(itertools.product([1, 2, 3, 8, 33, 125], [2, 3, 4, 8])))
def test_mg_cyclic_map_tiles_every_column_once(mB, P):
    tabs = [mg_tables(mB, P, r) for r in range(P)]
    assert sorted(I for own, _, _ in tabs for I in own) == list(range(mB))          # every block row has one owner
    for Bmin in range(mB):
        covered = set()
        for r in range(P):
            own, pref, lfirst = tabs[r]
            nl, l0 = len(own), lfirst[Bmin]
            strips = mg_F(pref, l0, Bmin, nl)
            assert strips == sum(SPT * (I - Bmin + 1) for I in own if I >= Bmin)
            worst = 0
            for u in range(strips):
                l, steps = mg_local_row(own, pref, l0, nl, Bmin, P, u)
                worst = max(worst, steps)
                k = u - mg_F(pref, l0, Bmin, l)
                b = own[l] - Bmin
                assert 0 <= k < SPT * (b + 1)
                key = (own[l], k)
                assert key not in covered
                covered.add(key)
            assert worst <= 3                                                         # the quadratic guess lands next to the answer
        assert len(covered) == SPT * (mB - Bmin) * (mB - Bmin + 1) // 2                 # the whole trailing triangle
    # balance: no rank streams more than one 128-row group above the mean
    for Bmin in range(0, mB, 7):
        per = []
        for r in range(P):
            own, pref, lfirst = tabs[r]
            per.append(mg_F(pref, lfirst[Bmin], Bmin, len(own)))
        assert max(per) - min(per) <= 2 * SPT * (mB - Bmin + 1)


# ---- panel gather of the sharded trailing update (comm.cu: cyc_rows_below / k_cyc_panel) ---------------------------
def cyc_rows_below(x, group, P, q):
    G, cyc, rem = x // group, (x // group) // P, (x // group) % P
    return cyc * group + (group if rem > q else 0) + (x % group if rem == q else 0)


@pytest.mark.parametrize("n,r0,P", [(1000, 0, 2), (1000, 64, 2), (1000, 448, 4), (5000, 1984, 8), (130, 128, 8), (700, 699, 3)])
def test_cyclic_row_packing_is_a_bijection(n, r0, P):
    group = 128
    lmax = max(cyc_rows_below(n, group, P, q) - cyc_rows_below(r0, group, P, q) for q in range(P))
    seen = {}
    for i in range(r0, n):
        q = (i // group) % P
        idx = cyc_rows_below(i, group, P, q) - cyc_rows_below(r0, group, P, q)
        assert 0 <= idx < lmax
        assert (q, idx) not in seen                      # no two rows of a rank share a slot
        seen[(q, idx)] = i
    for q in range(P):                                    # the slots of a rank are dense and in row order
        mine = sorted(idx for (qq, idx) in seen if qq == q)
        assert mine == list(range(len(mine)))
        rows = [seen[(q, k)] for k in mine]
        assert rows == sorted(rows)


# ---- sharded trailing update (gemm_ozaki.cu: ozaki_syr2k_lower_sub) vs kernel ownership (sytrd_mg.cuh) ---------------
@pytest.mark.parametrize("n,P", [(700, 2), (1500, 2), (2700, 4), (7954, 8), (4097, 3)])
def test_sharded_update_rows_cover_what_the_kernel_reads(n, P):
    """For every panel start r0 (multiples of 64) the 128-row tile rows a rank updates are exactly the block rows the
    panel kernel lets it read (owner of 64-row block row I = (I // 2) % P), and every trailing row has one updater."""
    BM, TILE = 128, 64
    for r0 in range(64, n, 64):
        nr = n - r0
        if nr < 384:                      # small tail: every rank updates everything
            continue
        lead = r0 % BM
        rows = nr + lead
        g0 = (r0 - lead) // BM
        ntile = (rows + BM - 1) // BM
        updater = {}
        for rank in range(P):
            for i in range(ntile):
                if (g0 + i) % P != rank:
                    continue
                for row in range(r0 - lead + i * BM, min(r0 - lead + (i + 1) * BM, n)):
                    if row >= r0:
                        assert row not in updater
                        updater[row] = rank
        assert sorted(updater) == list(range(r0, n))
        for row, rank in updater.items():
            assert ((row // TILE) >> 1) % P == rank


@pytest.mark.parametrize("n,P", [(64, 2), (700, 2), (2700, 4), (7954, 8), (65535, 8), (4097, 3)])
def test_library_mg_tables_equal_the_restatement(n, P):
    """The C++ table builder of the kernel (sytrd_mg.cuh: mg_tables_host, through rsoptsc_host_mg_tables; no GPU needed)
    against mg_tables above, which the map tests of this file are written on."""
    import ctypes as C

    import numpy as np
    from rsoptsc_b200 import _lib
    lib = _lib.load()
    mB = (n + 63) // 64
    for rank in range(P):
        own = np.zeros(mB, dtype=np.int32)
        pref = np.zeros(mB + 1, dtype=np.int64)
        lfirst = np.zeros(mB + 1, dtype=np.int32)
        nl = C.c_int32(0)
        st = lib.rsoptsc_host_mg_tables(n, P, rank, own.ctypes.data_as(_lib.c_int32_p), C.byref(nl),
                                        pref.ctypes.data_as(_lib.c_int64_p), lfirst.ctypes.data_as(_lib.c_int32_p))
        assert st == 0, lib.rsoptsc_last_error()
        o, p, lf = mg_tables(mB, P, rank)
        assert nl.value == len(o)
        assert own[:nl.value].tolist() == o
        assert pref[:nl.value + 1].tolist() == p
        assert lfirst.tolist() == lf
