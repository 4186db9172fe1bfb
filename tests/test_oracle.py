"""The oracle (oracle/blis_oracle.c) against the reference's own answers.

CPU only.  Two sources pin it:
  * golden fixtures in tests/golden/ generated from the real reference build
    (always available),
  * the real reference library oracle/_ref/libblis_ref.so when present (it is in
    the build container and on the GPU box).
Index arithmetic and packing must be bit-exact; gemm/trsm must be bit-exact on
power-of-two / integer inputs and within the testsuite-derived tolerance
otherwise (testsuite/src/test_gemm.c:44-47, test_trsm.c:44-47).
"""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

import gen
import make_golden as G
from refblis import DT, LEFT, LOWER, TRANSPOSE
from util import TOL, estr, rel_err

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def index_gold():
    return json.loads((GOLD / "index_arith.json").read_text())


def test_determine_blocksize_golden(oracle, index_gold):
    for bw, dim, b_alg, b_max, seq in index_gold["determine_blocksize"]:
        got, i = [], 0
        while i < dim:
            b = oracle.determine_blocksize(bw, i, dim, b_alg, b_max)
            assert b > 0
            got.append(b); i += b
        assert got == seq, (bw, dim, b_alg, b_max)


def test_thread_range_sub_golden(oracle, index_gold):
    for n_way, n, bf, low, ranges in index_gold["thread_range_sub"]:
        got = [list(oracle.thread_range_sub(w, n_way, n, bf, low)) for w in range(n_way)]
        assert got == ranges, (n_way, n, bf, low)
        # partition property: contiguous cover of [0, n)
        assert got[0][0] == 0 and got[-1][1] == n
        assert all(got[i][1] == got[i + 1][0] for i in range(n_way - 1))


def test_thread_partition_2x2_golden(oracle, index_gold):
    for nt, w1, w2, res in index_gold["thread_partition_2x2"]:
        got = list(oracle.thread_partition_2x2(nt, w1, w2))
        assert got == res and got[0] * got[1] == nt, (nt, w1, w2)


def test_index_arith_live_reference(oracle, ref):
    db, tr, p2 = G.index_cases()
    for bw, dim, b_alg, b_max in db[::7]:
        for i in (0, dim // 3, dim - 1):
            assert oracle.determine_blocksize(bw, i, dim, b_alg, b_max) == ref.determine_blocksize(bw, i, dim, b_alg, b_max)
    for n_way, n, bf, low in tr[::5]:
        for w in range(n_way):
            assert oracle.thread_range_sub(w, n_way, n, bf, low) == ref.thread_range_sub(w, n_way, n, bf, low)
    for nt, w1, w2 in p2[::3]:
        assert oracle.thread_partition_2x2(nt, w1, w2) == ref.thread_partition_2x2(nt, w1, w2)


def test_packm_micropanels_bit_exact(oracle):
    """oracle packm_struc_cxk == bli_??packm_struc_cxk, byte for byte (incl. the
    zero padding, unit/inverted diagonal, identity extension, zeroed unstored part)."""
    gold = np.load(GOLD / "packm.npz")
    cases = G.packm_cases()
    assert len(cases) == len(gold.files)
    for idx, cs in enumerate(cases):
        ch, tri, uplo, unit, conj, inv, pd, pl, pdm, plm, pdo, plo, kappa, order = cs
        src = gen.matrix(ch, pd, pl, 17 * idx + 5, "frac", order, pad=2)
        kap = np.array([kappa], dtype=src.dtype)
        p = np.full(pdm * plm, 777.0, dtype=src.dtype)
        rs, cs_ = estr(src)
        getattr(oracle.lib, f"orc_{ch}packm_struc_cxk")(tri, uplo, unit, conj, inv, pd, pl, pdm, plm, pdo, plo,
                                                        kap.ctypes.data, src.ctypes.data, rs, cs_, p.ctypes.data, pdm)
        want = gold[f"p{idx}"]
        if ch in "cz" and tri and inv and not unit:
            # The complex reciprocal (bli_tinverts) is the one packm step that is not a copy or a
            # single multiply; the reference compiles its ref kernels with -funsafe-math-optimizations
            # -ffp-contract=fast (config/*/make_defs.mk, CRVECFLAGS), so its last bit is not defined
            # by the source.  Everything except the inverted diagonal entries must still be identical
            # bytes, and those entries must agree to 4 ulp.
            eps = np.finfo(np.float32 if ch == "c" else np.float64).eps
            neq = p != want
            assert np.abs(p - want).max() <= 4 * eps * np.abs(want).max(), f"packm case {idx}: {cs}"
            assert neq.sum() <= pd, f"packm case {idx}: more than the diagonal differs"
        else:
            assert p.tobytes() == want.tobytes(), f"packm case {idx}: {cs}"


def _exact(kind):
    return kind in ("pow2", "ints")


def test_gemm_vs_golden(oracle):
    gold = np.load(GOLD / "gemm.npz")
    for idx, cs in enumerate(G.gemm_cases()):
        ch, kind = cs[0], cs[1]
        a, b, c = G.gemm_inputs(cs, idx)
        oracle.gemm(cs[5], cs[6], cs[10], a, b, cs[11], c)
        want = gold[f"c{idx}"]
        if _exact(kind):
            assert np.ascontiguousarray(c).tobytes() == want.tobytes(), f"gemm case {idx} {cs} not bit-exact"
        else:
            assert rel_err(c, want) <= TOL[ch], f"gemm case {idx} {cs}: {rel_err(c, want)}"


def test_trsm_vs_golden(oracle):
    gold = np.load(GOLD / "trsm.npz")
    for idx, cs in enumerate(G.trsm_cases()):
        ch, kind = cs[0], cs[1]
        a, b = G.trsm_inputs(cs, idx)
        oracle.trsm(cs[4], cs[5], cs[6], cs[7], cs[10], a, b)
        want = gold[f"x{idx}"]
        if _exact(kind):
            assert np.array_equal(np.ascontiguousarray(b), want), f"trsm case {idx} {cs} not bit-exact"
        else:
            assert rel_err(b, want) <= 20 * TOL[ch], f"trsm case {idx} {cs}: {rel_err(b, want)}"


def _tri_masks(m, lower):
    unstored = np.triu(np.ones((m, m), bool), 1) if lower else np.tril(np.ones((m, m), bool), -1)
    return ~unstored, unstored


def test_hemm_symm_trmm_vs_golden(oracle):
    """hemm / symm / trmm3 / trmm: the restatement against reference outputs; the unstored triangle of A is NaN, so a
    finite result also proves that it is never read."""
    gold = np.load(GOLD / "strucmm.npz")
    for idx, cs in enumerate(G.strucmm_cases()):
        ch, kind = cs[0], cs[2]
        a, b, c = G.strucmm_inputs(cs, idx)
        out = G.strucmm_run(oracle, cs, a, b, c)
        want = gold[f"c{idx}"]
        if _exact(kind):
            assert np.array_equal(np.ascontiguousarray(out), want), f"strucmm case {idx} {cs} not bit-exact"
        else:
            assert rel_err(out, want) <= TOL[ch], f"strucmm case {idx} {cs}: {rel_err(out, want)}"


def test_gemm_mixed_datatype_vs_golden(oracle):
    """All 4 x 4 x 4 storage datatype combinations x both computation precisions: the restatement of
    bli_gemm_cntl.c's mixed-domain / mixed-precision rules against reference outputs (bit-exact on power-of-two data;
    one reference artefact, a denormal left in a real C by the real-only packing path, is tolerated below 1e-300)."""
    from refblis import oracle_gemm_md
    gold = np.load(GOLD / "gemm_md.npz")
    for idx, cs in enumerate(G.gemm_md_cases()):
        a, b, c = G.gemm_md_inputs(cs, idx)
        oracle_gemm_md(oracle, cs[8], cs[9], cs[11], a, b, cs[12], c, cs[3])
        want = gold[f"c{idx}"]
        if _exact(cs[4]):
            assert float(np.abs(np.ascontiguousarray(c) - want).max()) < 1e-300, f"gemm_md case {idx} {cs} not bit-exact"
        else:
            assert rel_err(c, want) <= G.md_tol(cs), f"gemm_md case {idx} {cs}: {rel_err(c, want)}"


def test_gemmt_family_vs_golden(oracle):
    """gemmt / syrk / herk / syr2k / her2k: the restatement against reference outputs.  The triangle of C that is
    not stored is NaN in the inputs and must come back untouched (the reference leaves it NaN too)."""
    gold = np.load(GOLD / "gemmt.npz")
    for idx, cs in enumerate(G.gemmt_cases()):
        ch, op, kind, m, uplo = cs[0], cs[1], cs[2], cs[3], cs[5]
        a, b, c = G.gemmt_inputs(cs, idx)
        G.gemmt_run(oracle, cs, a, b, c)
        want = gold[f"c{idx}"]
        stored, unstored = _tri_masks(m, uplo == G.LOWER)
        assert np.isnan(np.abs(c[unstored])).all(), f"gemmt case {idx} {cs}: unstored triangle written"
        if _exact(kind):
            assert np.array_equal(c[stored], want[stored]), f"gemmt case {idx} {cs} not bit-exact"
        else:
            assert rel_err(c[stored], want[stored]) <= TOL[ch], f"gemmt case {idx} {cs}: {rel_err(c[stored], want[stored])}"
        if op in ("herk", "her2k") and ch in "cz":
            assert (np.diag(c).imag == 0).all()


@pytest.mark.parametrize("ch", list("sdcz"))
def test_gemm_live_reference_other_blocksizes(oracle, ref, ch):
    """Same algorithm under the haswell-like blocksizes (incl. row preference):
    results stay within tolerance of the real reference."""
    bs = {"s": (6, 16, 144, 256, 4080), "d": (6, 8, 72, 256, 4080), "c": (3, 8, 144, 256, 4080), "z": (3, 4, 72, 256, 4080)}[ch]
    oracle.set_blksz(ch, *bs, row_pref=1)
    try:
        for idx, (m, n, k) in enumerate(((150, 90, 40), (7, 300, 260), (97, 1, 513))):
            for oc in "cr":
                a = gen.matrix(ch, m, k, 900 + idx, "frac", "c"); b = gen.matrix(ch, k, n, 950 + idx, "frac", "r")
                c1 = gen.matrix(ch, m, n, 990 + idx, "frac", oc); c2 = c1.copy(order="K")
                oracle.gemm(0, 0, 2.0, a, b, 1.2, c1); ref.gemm(0, 0, 2.0, a, b, 1.2, c2)
                assert rel_err(c1, c2) <= TOL[ch]
    finally:
        d = {"s": (4, 16, 256, 256, 4096), "d": (4, 8, 128, 256, 4096), "c": (4, 8, 128, 256, 4096), "z": (4, 4, 64, 256, 4096)}[ch]
        oracle.set_blksz(ch, *d, row_pref=0)


@pytest.mark.parametrize("ch", list("sdcz"))
def test_trsm_live_reference_multiblock(oracle, ref, ch):
    """m spanning several KC blocks and ragged MR/NR edges, all side/uplo combos."""
    al = (2.0 + 0.3j) if ch in "cz" else 2.0
    idx = 0
    for (m, n) in ((300, 70), (257, 130)):
        for side in (0, 1):
            for uplo in (0x60, 0xC0):
                for tr in (0, 8, 0x18):
                    idx += 1
                    ma = m if side == LEFT else n
                    a = gen.triangular(ch, ma, 2000 + idx, "frac", "c")
                    b1 = gen.matrix(ch, m, n, 3000 + idx, "frac", "c"); b2 = b1.copy(order="K")
                    oracle.trsm(side, uplo, tr, 0, al, a, b1); ref.trsm(side, uplo, tr, 0, al, a, b2)
                    assert rel_err(b1, b2) <= 20 * TOL[ch], (m, n, side, uplo, tr)
