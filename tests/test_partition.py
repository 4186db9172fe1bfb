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
