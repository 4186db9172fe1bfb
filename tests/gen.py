"""Deterministic, platform-independent synthetic inputs shared by the golden
generator (tests/golden/make_golden.py) and the tests.

Values come from integer hashes, not from an RNG stream, so the same matrices
are rebuilt bit-for-bit wherever the tests run (this container, the GPU box):

* `frac`  -- multiples of 1/128 in [-1, 1): exact in fp32, the analogue of the
             testsuite's uniform [-1,1] inputs (testsuite/src/test_libblis.c:2529-2565)
* `pow2`  -- 0 or +-2^-e, e in [0,4]: the testsuite's "powers of two in a narrow
             precision range" mode (bli_randnp2s, frame/include/bli_cast_macro_defs.h:476-524);
             products and moderate sums are exact, so ANY summation order must
             give identical bits
* `ints`  -- small integers (for exact triangular solves)
"""
from __future__ import annotations

import numpy as np

NP_DT = {"s": np.float32, "d": np.float64, "c": np.complex64, "z": np.complex128}


def _hash(i, j, seed):
    x = (i.astype(np.uint64) * np.uint64(2654435761) + j.astype(np.uint64) * np.uint64(40503)
         + np.uint64(seed) * np.uint64(2246822519) + np.uint64(12345))
    x ^= x >> np.uint64(13)
    x *= np.uint64(0x9E3779B97F4A7C15)
    x ^= x >> np.uint64(29)
    return x


def _real(m, n, seed, kind):
    i, j = np.meshgrid(np.arange(m), np.arange(n), indexing="ij")
    h = _hash(i, j, seed)
    if kind == "frac":
        return ((h % np.uint64(256)).astype(np.float64) - 128.0) / 128.0
    if kind == "pow2":
        e = (h % np.uint64(5)).astype(np.float64)
        sgn = np.where((h >> np.uint64(8)) % np.uint64(2) == 0, 1.0, -1.0)
        zero = (h >> np.uint64(16)) % np.uint64(7) == 0
        return np.where(zero, 0.0, sgn * np.exp2(-e))
    if kind == "ints":
        return (h % np.uint64(5)).astype(np.float64) - 2.0
    raise ValueError(kind)


def matrix(ch: str, m: int, n: int, seed: int, kind: str = "frac", order: str = "c", pad: int = 0) -> np.ndarray:
    """m x n matrix of datatype ch ('s','d','c','z'); order 'c' column-major,
    'r' row-major, 'g' general stride (rs=2-ish, cs padded); pad adds to the
    leading dimension."""
    dt = NP_DT[ch]
    v = _real(m, n, seed, kind)
    if ch in "cz":
        v = v + 1j * _real(m, n, seed + 1000003, kind)
    v = v.astype(dt)
    if order == "c":
        buf = np.zeros((n, m + pad), dtype=dt)
        out = buf.T[:m, :]
    elif order == "r":
        buf = np.zeros((m, n + pad), dtype=dt)
        out = buf[:, :n]
    else:
        buf = np.zeros((2 * m + 1, 3 * n + 2 + pad), dtype=dt)
        out = buf[1::2, 2::3][:m, :n]
    out[...] = v
    return out


def triangular(ch: str, m: int, seed: int, kind: str = "frac", order: str = "c") -> np.ndarray:
    """Full m x m matrix whose diagonal is made safely non-singular; the side
    that trsm must NOT read is filled with NaN by the caller when wanted."""
    a = matrix(ch, m, m, seed, kind, order)
    d = np.arange(m)
    if kind == "ints":
        a[d, d] = np.where(_hash(d, d, seed + 7) % np.uint64(2) == 0, 1.0, -1.0).astype(a.dtype)
    elif kind == "pow2":
        a[d, d] = (np.where(_hash(d, d, seed + 7) % np.uint64(2) == 0, 1.0, -1.0)
                   * np.exp2((_hash(d, d, seed + 9) % np.uint64(3)).astype(np.float64))).astype(a.dtype)
    else:
        # testsuite: random then diag += 2.0 (testsuite/src/test_libblis.c:2583-2589); scaled to keep
        # the solve well conditioned at any m
        a[...] = a / max(1.0, np.sqrt(m) / 2)
        a[d, d] += 2.0
    return a


def poison_unstored(a: np.ndarray, uplo_lower: bool) -> np.ndarray:
    """NaN-fill the triangle trsm must not read (catches any stray access)."""
    m = a.shape[0]
    mask = np.triu(np.ones((m, m), dtype=bool), 1) if uplo_lower else np.tril(np.ones((m, m), dtype=bool), -1)
    a[mask] = np.nan
    return a
