"""blis_b200 -- B200-native level-3 (gemm/trsm) engine behind BLIS's API.

The product is the C-ABI library `libblis_b200.so` (include/blis_b200.h) built
from `blis_b200/csrc/*.cu` for sm_100a; this package is the thin host-side
mirror of the reference's typed/object/BLAS APIs used by tests and bench.py.
"""
from ._lib import (BLIS_CONJ_NO_TRANSPOSE, BLIS_CONJ_TRANSPOSE, BLIS_DCOMPLEX, BLIS_DOUBLE,
                   BLIS_FAILURE, BLIS_FLOAT, BLIS_LEFT, BLIS_LOWER, BLIS_NO_TRANSPOSE,
                   BLIS_NONUNIT_DIAG, BLIS_RIGHT, BLIS_SCOMPLEX, BLIS_SUCCESS, BLIS_TRANSPOSE,
                   BLIS_UNIT_DIAG, BLIS_UPPER, EngineError)

__all__ = [n for n in dir() if n.startswith("BLIS_")] + ["EngineError"]
