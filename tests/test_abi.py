"""The C-ABI boundary without a GPU: the library loads, exports every symbol
include/blis_b200.h declares, reports the registered tile shapes, and fails
LOUDLY (no CPU fallback) when no device is present.  Also the host-side
argument checking of the API mirror (frame/compat/check, frame/3/bli_l3_check.c)."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    h = (ROOT / "include" / "blis_b200.h").read_text()
    names = set(re.findall(r"\b(b200_[a-z_0-9]+)\s*\(", h))
    for ch in "sdcz":                      # B200_DECL_GEMM / B200_DECL_TRSM expansions
        names.add(f"b200_{ch}gemm"); names.add(f"b200_{ch}trsm")
    names.discard("b200_")                 # the macro body's token paste
    return {n for n in names if not n.endswith("_")}


def test_library_exports_every_declared_symbol():
    from blis_b200 import _lib
    lib = _lib.load()
    decl = _declared_symbols()
    assert {"b200_gemm", "b200_trsm", "b200_dgemm", "b200_ztrsm", "b200_init", "b200_blksz"} <= decl
    missing = [n for n in sorted(decl) if not hasattr(lib, n)]
    assert not missing, f"declared in include/blis_b200.h but not exported: {missing}"
    assert set(_lib.EXPORTS) <= decl


def test_enum_values_match_reference_header():
    """include/blis_b200.h restates BLIS's enum values; compare with the reference when mounted."""
    from blis_b200 import _lib
    hdr = Path("/root/reference/frame/include/bli_type_defs.h")
    assert (_lib.BLIS_NO_TRANSPOSE, _lib.BLIS_TRANSPOSE, _lib.BLIS_CONJ_NO_TRANSPOSE, _lib.BLIS_CONJ_TRANSPOSE) == (0, 8, 16, 24)
    assert (_lib.BLIS_UPPER, _lib.BLIS_LOWER, _lib.BLIS_UNIT_DIAG) == (0x60, 0xC0, 0x100)
    assert (_lib.BLIS_SUCCESS, _lib.BLIS_FAILURE) == (-1, -2)
    if hdr.exists():
        t = hdr.read_text()
        assert "BLIS_SUCCESS                               = ( -1)" in t or re.search(r"BLIS_SUCCESS\s*=\s*\(\s*-1\)", t)


def test_blocksizes_satisfy_blis_registration_invariants():
    """bli_gks_register_cntx aborts unless MC%MR == NC%NR == 0 etc.
    (frame/base/bli_gks.c:248-271); the tile shapes we hand to bli_cntx_set_blkszs must pass."""
    from blis_b200 import api
    for dt in (torch.float32, torch.float64, torch.complex64, torch.complex128):
        mr, nr, mc, kc, nc = (api.blksz(dt, w) for w in ("MR", "NR", "MC", "KC", "NC"))
        assert min(mr, nr, mc, kc, nc) > 0
        assert mc % mr == 0 and nc % nr == 0 and mc % nr == 0 and nc % mr == 0
        assert mr % 2 == 0 or nr % 2 == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    from blis_b200 import EngineError, api
    a = torch.zeros(4, 4, dtype=torch.float64)
    with pytest.raises(EngineError, match="no CUDA device|no CPU fallback"):
        api.bli_dgemm(0, 0, 4, 4, 4, 1.0, a, 4, 1, a, 4, 1, 0.0, a.clone(), 4, 1)
    with pytest.raises(EngineError):
        api.measure_peak("dmma", 10)


def test_blas_layer_argument_checks():
    """xerbla-style parameter checks of ?gemm_/?trsm_ (dblat3 'TESTS OF ERROR-EXITS')."""
    from blis_b200 import api
    z = torch.zeros(4, 4, dtype=torch.float64)
    with pytest.raises(ValueError, match="parameter 1"):
        api.dgemm_("X", "N", 4, 4, 4, 1.0, z, 4, z, 4, 0.0, z, 4)
    with pytest.raises(ValueError, match="parameter 3"):
        api.dgemm_("N", "N", -1, 4, 4, 1.0, z, 4, z, 4, 0.0, z, 4)
    with pytest.raises(ValueError, match="parameter 8"):
        api.dgemm_("N", "N", 4, 4, 4, 1.0, z, 3, z, 4, 0.0, z, 4)
    with pytest.raises(ValueError, match="parameter 13"):
        api.dgemm_("T", "N", 4, 4, 4, 1.0, z, 4, z, 4, 0.0, z, 2)
    with pytest.raises(ValueError, match="parameter 2"):
        api.dtrsm_("L", "Q", "N", "N", 4, 4, 1.0, z, 4, z, 4)
    with pytest.raises(ValueError, match="parameter 9"):
        api.dtrsm_("R", "L", "N", "N", 2, 4, 1.0, z, 3, z, 4)


def test_object_api_checks():
    from blis_b200 import BLIS_LEFT, BLIS_LOWER, EngineError, api
    a = api.Obj(torch.zeros(4, 5, dtype=torch.float64)); b = api.Obj(torch.zeros(6, 3, dtype=torch.float64))
    c = api.Obj(torch.zeros(4, 3, dtype=torch.float64))
    with pytest.raises(EngineError, match="non-conformal"):
        api.bli_gemm(1.0, a, b, 0.0, c)
    with pytest.raises(EngineError, match="no CPU fallback"):      # mixed datatypes are dispatched to b200_gemm_md: needs the GPU
        api.bli_gemm(1.0, a, api.Obj(torch.zeros(5, 3, dtype=torch.float32)), 0.0, c)
    t = api.Obj(torch.zeros(4, 4, dtype=torch.float64))
    with pytest.raises(EngineError, match="triangular"):
        api.bli_trsm(BLIS_LEFT, 1.0, t, c)
    api.bli_obj_set_uplo(BLIS_LOWER, t)
    with pytest.raises(EngineError, match="non-conformal"):
        api.bli_trsm(BLIS_LEFT, 1.0, t, api.Obj(torch.zeros(5, 3, dtype=torch.float64)))


def test_every_entry_point_is_in_the_binding_table():
    """INTEGRATION.md's table names the reference interface behind every symbol include/blis_b200.h declares."""
    import re
    hdr = (ROOT / "include" / "blis_b200.h").read_text()
    doc = (ROOT / "INTEGRATION.md").read_text()
    syms = sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(syms) >= 25
    typed = re.compile(r"b200_[sdcz](gemm|trsm)$")           # listed as b200_?gemm / b200_?trsm
    missing = [s_ for s_ in syms if s_ not in doc and not typed.match(s_)]
    assert not missing, missing
