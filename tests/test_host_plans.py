"""Host-side schedule arithmetic of the engine, checked without a GPU through the C ABI:

* b200_splitk_plan  -- when mid-size dgemm cuts the tiles of its partial last wave into k chunks (csrc/gemm_d.cu);
* b200_trsm_upload_plan -- the order in which a host-resident triangular A travels under the solve (csrc/host_trsm.cuh):
  its pieces must tile the stored triangle exactly (plus the unstored half of the small diagonal squares) and provide
  exactly one event per launch of the recursive solve, in the recursion's order;
* b200_trsm_rowblock_plan -- the same for the row-block (left-looking) pipeline of tall systems: one update gemm per block
  row reading A[j, 0:j], then the recursion inside the diagonal block.
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
    l.b200_trsm_rowblock_plan.restype = C.c_int
    l.b200_trsm_rowblock_plan.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_int64), C.c_int]
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


def rowblock_launches(m, leaf, upper, blocks):
    """The launches of trsm_host_rowpipe (csrc/host_trsm.cuh) as regions of A in launch order: per block row the update
    gemm over everything already solved, then trsm_rec on the diagonal block."""
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
    for r0, r1 in blocks:
        if not upper and r0 > 0:
            out.append((r0, r1, 0, r0))
        if upper and r1 < m:
            out.append((r0, r1, r1, m))
        rec(r0, r1 - r0)
    return out


@pytest.mark.parametrize("upper", [False, True])
@pytest.mark.parametrize("m,leaf,rb", [(4608, 256, 1024), (32768, 256, 0), (16384, 256, 0), (5000, 256, 1280), (4100, 64, 1000), (8192, 32, 0),
                                       (6144, 256, 4096), (9999, 256, 0)])
def test_trsm_rowblock_plan_matches_the_solve(lib, m, leaf, upper, rb):
    cap = 8 * (m // leaf + 2)
    buf = (C.c_int64 * (5 * cap))()
    n = lib.b200_trsm_rowblock_plan(m, 8192, leaf, int(upper), rb, buf, cap)
    assert 0 < n <= cap
    plan = np.ctypeslib.as_array(buf)[:5 * n].reshape(n, 5).copy()
    # block rows, in processing order, are the B pieces (c0 = c1 = -1, no launch waits for them: the stream's order does)
    blocks = [(int(r0), int(r1)) for r0, r1, c0, c1, nl in plan if c0 < 0]
    assert all(nl == 0 for r0, r1, c0, c1, nl in plan if c0 < 0) and len(blocks) >= 2
    order = blocks[::-1] if upper else blocks
    assert order[0][0] == 0 and order[-1][1] == m and all(a[1] == b[0] for a, b in zip(order, order[1:]))     # they tile [0, m)
    assert all(r0 % leaf == 0 for r0, _ in blocks)                                                            # on leaf boundaries
    mr = -(-m // leaf) * leaf                                 # sizes are counted in whole leaves: the block that ends at row m is cut there
    sizes = [(mr if r1 == m else r1) - r0 for r0, r1 in blocks]
    if rb:
        rows = -(-rb // leaf) * leaf                      # explicit: uniform (the last one takes a remainder below half a block)
        assert all(s == rows for s in sizes[:-1]) and sizes[-1] < 1.5 * rows + leaf
    else:
        rows = -(-max(256, -(-(m // 32) // 256) * 256) // leaf) * leaf      # the engine's choice: m/32 in whole 256 rows
        assert all(s == rows for s in sizes[:-1]) and sizes[-1] < 1.5 * rows + leaf
    pieces = plan[plan[:, 2] >= 0]
    launches = rowblock_launches(m, leaf, upper, blocks)
    assert int(pieces[:, 4].sum()) == len(launches)
    assert len(launches) == sum(2 * (-(-(r1 - r0) // leaf)) - 1 for r0, r1 in blocks) + len(blocks) - 1
    owner = np.repeat(np.arange(len(pieces)), pieces[:, 4])
    for (r0, r1, c0, c1), p in zip(launches, owner):
        pr0, pr1, pc0, pc1, _ = pieces[p]
        assert pr0 <= r0 and r1 <= pr1 and pc0 <= c0 and c1 <= pc1, (m, leaf, upper, (r0, r1, c0, c1), pieces[p])
    # every piece of A travels after the rows of B it is used on, and reads X only where it was solved by then: its rows and
    # its columns lie in block rows whose B has already gone up
    seen = set()
    for r0, r1, c0, c1, nl in plan:
        if c0 < 0:
            seen.update(range(int(r0) // leaf, -(-int(r1) // leaf)))
        else:
            assert all(x in seen for x in (int(r0) // leaf, (int(r1) - 1) // leaf, int(c0) // leaf, (int(c1) - 1) // leaf))
    if m <= 10000:
        cover = np.zeros((m, m), dtype=np.int8)
        for r0, r1, c0, c1, _ in pieces:
            cover[r0:r1, c0:c1] += 1
        assert cover.max() == 1
        tri = np.triu(np.ones((m, m), dtype=bool)) if upper else np.tril(np.ones((m, m), dtype=bool))
        assert cover[tri].min() == 1
        assert int(cover.sum()) - int(tri.sum()) <= m * max(leaf, 1024) // 2


def test_trsm_rowblock_plan_declines_short_systems(lib):
    buf = (C.c_int64 * 50)()
    assert lib.b200_trsm_rowblock_plan(2048, 8192, 256, 0, 0, buf, 10) == 0       # below trsm_host_rb_min_m
    assert lib.b200_trsm_rowblock_plan(32768, 512, 256, 0, 0, buf, 10) == 0       # too few right-hand sides
    assert lib.b200_trsm_rowblock_plan(32768, 8192, 256, 0, -1, buf, 10) == 0     # switched off
    assert lib.b200_trsm_rowblock_plan(2048, 8192, 256, 0, 4096, buf, 10) == 0    # one block is no pipeline


@pytest.mark.parametrize("upper", [False, True])
@pytest.mark.parametrize("m,leaf,rb", [(1576, 64, 512), (1000, 32, 300), (2048, 256, 512)])
def test_trsm_rowblock_plan_is_sufficient_for_the_arithmetic(lib, m, leaf, upper, rb):
    """The row-block pipeline replayed on the CPU exactly as csrc/host_trsm.cuh queues it: the device images of A and B
    start as NaN, the plan's pieces are copied in the plan's order, and after each piece of A the launches that wait for
    it run (update gemm of the block row, then the recursion's leaves and updates).  A launch that reads a byte that has
    not travelled yet, or a row of X that is not final yet, poisons the result; the result must equal the solve of the
    whole system.  alpha is applied exactly once per row, by the first operation that touches it."""
    n, alpha = 24, 2.0
    cap = 8 * (m // leaf + 2)
    buf = (C.c_int64 * (5 * cap))()
    # m < trsm_host_rb_min_m: an explicit block size still cuts the system (the engine's own choice would decline)
    cnt = lib.b200_trsm_rowblock_plan(m, 8192, leaf, int(upper), rb, buf, cap)
    assert 0 < cnt <= cap
    plan = np.ctypeslib.as_array(buf)[:5 * cnt].reshape(cnt, 5).copy()
    blocks = [(int(r0), int(r1)) for r0, r1, c0, c1, nl in plan if c0 < 0]
    launches = iter(rowblock_launches(m, leaf, upper, blocks))
    rng = np.random.default_rng(m + leaf + upper)
    a = rng.uniform(-1, 1, (m, m)) / np.sqrt(m); a[np.arange(m), np.arange(m)] += 2.0
    a = np.triu(a) if upper else np.tril(a)
    b = rng.uniform(-1, 1, (m, n))
    da, db = np.full((m, m), np.nan), np.full((m, n), np.nan)
    scaled = np.zeros(m, dtype=bool)

    def once(r0, r1):                       # alpha for rows that nobody has touched yet, 1 afterwards
        assert scaled[r0:r1].all() or not scaled[r0:r1].any()
        f = 1.0 if scaled[r0] else alpha
        scaled[r0:r1] = True
        return f

    for r0, r1, c0, c1, nl in plan:
        if c0 < 0:
            db[r0:r1] = b[r0:r1]
            continue
        da[r0:r1, c0:c1] = np.where(np.isnan(a[r0:r1, c0:c1]), 0.0, a[r0:r1, c0:c1])
        for _ in range(nl):
            lr0, lr1, lc0, lc1 = next(launches)
            if (lr0, lr1) == (lc0, lc1):    # leaf: triangular solve of a diagonal block, only its stored triangle is read
                t = np.triu(da[lr0:lr1, lr0:lr1]) if upper else np.tril(da[lr0:lr1, lr0:lr1])
                db[lr0:lr1] = np.linalg.solve(t, once(lr0, lr1) * db[lr0:lr1])
            else:                           # update: B[rows] := alpha_once * B[rows] - A[rows, cols] X[cols]
                assert scaled[lc0:lc1].all()
                db[lr0:lr1] = once(lr0, lr1) * db[lr0:lr1] - da[lr0:lr1, lc0:lc1] @ db[lc0:lc1]
    assert next(launches, None) is None
    want = np.linalg.solve(a, alpha * b)
    assert not np.isnan(db).any()
    assert np.abs(db - want).max() <= 1e-11 * max(1.0, np.abs(want).max())
