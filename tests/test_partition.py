"""blis_b200.partition (product-side host logic) must reproduce the reference's
index arithmetic bit for bit: checked against the golden vectors from the real library."""
import json
from pathlib import Path

from blis_b200 import partition as P

GOLD = json.loads((Path(__file__).resolve().parent / "golden" / "index_arith.json").read_text())


def test_determine_blocksize():
    for bw, dim, b_alg, b_max, seq in GOLD["determine_blocksize"]:
        got, i = [], 0
        while i < dim:
            b = P.determine_blocksize(bool(bw), i, dim, b_alg, b_max); got.append(b); i += b
        assert got == seq


def test_thread_range_sub():
    for n_way, n, bf, low, ranges in GOLD["thread_range_sub"]:
        assert [list(P.thread_range_sub(w, n_way, n, bf, bool(low))) for w in range(n_way)] == ranges


def test_thread_partition_2x2():
    for nt, w1, w2, res in GOLD["thread_partition_2x2"]:
        assert list(P.thread_partition_2x2(nt, w1, w2)) == res


# The same arithmetic behind the C ABI (blis_b200/csrc/host_dist.cuh: what b200_dist_gemm / b200_dist_trsm split by);
# host-only entry points, so they run without a GPU.
def test_c_abi_range_sub_matches_the_reference():
    from blis_b200 import api
    for n_way, n, bf, low, ranges in GOLD["thread_range_sub"]:
        assert [list(api.range_sub(w, n_way, n, bf, bool(low))) for w in range(n_way)] == ranges


def test_c_abi_partition_2x2_matches_the_reference():
    from blis_b200 import api
    for nt, w1, w2, res in GOLD["thread_partition_2x2"]:
        assert list(api.partition_2x2(nt, w1, w2)) == res


def test_c_abi_dist_plan_equals_the_python_plan():
    """b200_dist_plan (grid, my block of C, panel ownership) against blis_b200.dist.SummaPlan for every rank of several
    worlds and ragged shapes."""
    import pytest
    from blis_b200 import api
    from blis_b200._lib import EngineError
    from blis_b200.dist import SummaPlan
    for world in (1, 2, 3, 4, 6, 8):
        for (m, n) in ((16384, 16384), (1000, 3001), (4097, 129), (32768, 65536)):
            for rank in range(world):
                try:
                    sp = SummaPlan(world, rank, m, n, 96 * 24, 96)
                except ValueError:
                    with pytest.raises(EngineError):
                        api.dist_plan(world, rank, m, n, 96 * 24, 96)
                    continue
                p = api.dist_plan(world, rank, m, n, 96 * 24, 96)
                assert (p.pr, p.pc, p.i, p.j, p.L, p.T, p.steps) == (sp.pr, sp.pc, sp.i, sp.j, sp.L, sp.T, sp.steps)
                assert (p.m0, p.m1, p.n0, p.n1) == (sp.m0, sp.m1, sp.n0, sp.n1)
                assert (p.na, p.nb) == (len(sp.a_panels()), len(sp.b_panels()))
