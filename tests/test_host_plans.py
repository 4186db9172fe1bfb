"""Host-side schedule arithmetic of the engine, checked without a GPU through the C ABI:

* b200_splitk_plan  -- when mid-size dgemm cuts the tiles of its partial last wave into k chunks (csrc/gemm_d.cu);
* b200_trsm_upload_plan -- the order in which a host-resident triangular A travels under the solve (csrc/host_trsm.cuh):
  its pieces must tile the stored triangle exactly (plus the unstored half of the small diagonal squares) and provide
  exactly one event per launch of the recursive solve, in the recursion's order.
"""
import ctypes as C

import numpy as np
import pytest

from blis_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    l = _lib.load()
    l.b200_splitk_plan.restype = C.c_int
    l.b200_splitk_plan.argtypes = [C.c_int64, C.c_int, C.c_int64, C.POINTER(C.c_int)]
    l.b200_trsm_upload_plan.restype = C.c_int
    l.b200_trsm_upload_plan.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_int]
    return l


def splitk(lib, tiles, grid, kt):
    full = C.c_int(-1)
    s = lib.b200_splitk_plan(tiles, grid, kt, C.byref(full))
    return s, full.value


def test_splitk_plan_known_cases(lib):
    G = 148
    # 2048^3: 256 tiles, 128 k steps -> 148 whole tiles + 108 x 4 chunks (432 units = 2.92 rounds of 1/4 tile)
    assert splitk(lib, 256, G, 128) == (4, 148)
    # 16384^3: 110.7 waves -- the tail is noise
    assert splitk(lib, 128 * 128, G, 1024)[0] == 0
    # whole waves, fewer tiles than one wave that already fill it, tiny k
    assert splitk(lib, 296, G, 128)[0] == 0
    assert splitk(lib, 144, G, 96)[0] == 0
    assert splitk(lib, 100, G, 80)[0] == 0                   # less than one wave: chunks only were measured slower
    assert splitk(lib, 256, G, 16)[0] == 0                   # 256 k: chunks would be shorter than 256 k


def test_splitk_plan_properties(lib):
    for grid in (148, 132, 296):
        for tiles in range(1, 1400, 7):
            for kt in (32, 64, 128, 300, 1024):
                s, full = splitk(lib, tiles, grid, kt)
                waves = -(-tiles // grid)
                if s == 0:
                    assert full == tiles
                    continue
                assert s in (2, 3, 4, 5, 6, 8)
                r = tiles - full
                assert full % grid == 0 and 0 < r < grid and waves <= 8
                assert kt // s >= 16                             # at least 256 k per chunk
                # the tail must get cheaper by at least 3 % of the launch, in the plan's own cost model
                tail = -(-r * s // grid) * (1.0 / s + 5.0 / kt)
                assert (1.0 - tail) / waves >= 0.03 - 1e-12
                # every chunk range is non-empty and they partition [0, kt)
                cuts = [kt * c // s for c in range(s + 1)]
                assert cuts[0] == 0 and cuts[-1] == kt and all(b > a for a, b in zip(cuts, cuts[1:]))


def upload_plan(lib, m, leaf, upper):
    cap = 4 * (m // leaf + 2)
    buf = (C.c_int64 * (5 * cap))()
    n = lib.b200_trsm_upload_plan(m, leaf, int(upper), buf, cap)
    assert 0 < n <= cap
    return np.ctypeslib.as_array(buf)[:5 * n].reshape(n, 5).copy()


def solve_launches(m, leaf, upper):
    """The launches of trsm_rec (csrc/host_trsm.cuh) as (rows, cols) regions of A in launch order."""
    out = []

    def rec(i0, mb):
        if mb <= leaf:
            out.append((i0, i0 + mb, i0, i0 + mb)); return
        nblk = -(-mb // leaf)
        m1 = ((nblk + 1) // 2) * leaf; m2 = mb - m1
        if not upper:
            rec(i0, m1); out.append((i0 + m1, i0 + mb, i0, i0 + m1)); rec(i0 + m1, m2)
        else:
            rec(i0 + m2, m1); out.append((i0, i0 + m2, i0 + m2, i0 + mb)); rec(i0, m2)
    rec(0, m)
    return out


@pytest.mark.parametrize("upper", [False, True])
@pytest.mark.parametrize("m,leaf", [(4608, 256), (32768, 256), (5000, 256), (4100, 64), (2048, 256), (1000, 256), (8192, 32)])
def test_trsm_upload_plan_matches_the_solve(lib, m, leaf, upper):
    plan = upload_plan(lib, m, leaf, upper)
    launches = solve_launches(m, leaf, upper)
    assert int(plan[:, 4].sum()) == len(launches) == 2 * (-(-m // leaf)) - 1
    # launch l waits for piece(l): the piece must contain everything the launch reads of A
    owner = np.repeat(np.arange(len(plan)), plan[:, 4])
    for (r0, r1, c0, c1), p in zip(launches, owner):
        pr0, pr1, pc0, pc1, _ = plan[p]
        assert pr0 <= r0 and r1 <= pr1 and pc0 <= c0 and c1 <= pc1, (m, leaf, upper, (r0, r1, c0, c1), plan[p])
    # pieces are disjoint and cover the stored triangle
    if m <= 8192:
        cover = np.zeros((m, m), dtype=np.int8)
        for r0, r1, c0, c1, _ in plan:
            cover[r0:r1, c0:c1] += 1
        assert cover.max() == 1
        tri = np.triu(np.ones((m, m), dtype=bool)) if upper else np.tril(np.ones((m, m), dtype=bool))
        assert cover[tri].min() == 1
        # what travels beyond the triangle is the unstored half of diagonal squares of at most max(leaf, 1024) rows
        extra = int(cover.sum()) - int(tri.sum())
        assert extra <= m * max(leaf, 1024) // 2
