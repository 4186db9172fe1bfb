"""ctypes access to the checkers (TEST INFRASTRUCTURE ONLY):

* `Oracle`  -- oracle/liboracle.so, the C restatement of the reference algorithm
* `RefBlis` -- oracle/_ref/libblis_ref.so, the real reference BLIS built from
               /root/reference by oracle/build_ref.py (present in the build
               container and shipped to the GPU box with the snapshot)

Nothing under blis_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "liboracle.so"
REF_SO = ROOT / "oracle" / "_ref" / "libblis_ref.so"

NO_TRANSPOSE, TRANSPOSE, CONJ_NO_TRANSPOSE, CONJ_TRANSPOSE = 0x00, 0x08, 0x10, 0x18
UPPER, LOWER, DENSE = 0x60, 0xC0, 0xE0
LEFT, RIGHT = 0, 1
NONUNIT_DIAG, UNIT_DIAG = 0x000, 0x100
CH = {np.dtype(np.float32): "s", np.dtype(np.float64): "d", np.dtype(np.complex64): "c", np.dtype(np.complex128): "z"}
DT = {"s": 0, "c": 1, "d": 2, "z": 3}
i64, vp, ci = C.c_int64, C.c_void_p, C.c_int


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


def _estr(a: np.ndarray):
    """(rs, cs) in elements of a 2-D numpy array."""
    return a.strides[0] // a.itemsize, a.strides[1] // a.itemsize


def _scalar(dtype, v):
    return np.array([v], dtype=dtype)


def build_oracle() -> Path:
    src = ROOT / "oracle" / "blis_oracle.c"
    inc = ROOT / "oracle" / "blis_oracle_t.inc"
    if not ORACLE_SO.exists() or ORACLE_SO.stat().st_mtime < max(src.stat().st_mtime, inc.stat().st_mtime):
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    return ORACLE_SO


class _GemmtFamily:
    """gemmt / syrk / herk / syr2k / her2k with the typed-API argument lists (frame/3/bli_l3_tapi.c:77-296), in place on
    numpy arrays with any strides.  `_PREFIX` is "orc_" (restatement) or "bli_" (the real reference)."""
    _PREFIX = ""

    def _two(self, name, uplo, transa, transb, alpha, a, b, beta, c):
        ch = CH[c.dtype]
        m = c.shape[0]
        k = a.shape[0] if (transa & TRANSPOSE) else a.shape[1]
        rdt = np.zeros(1, c.dtype).real.dtype
        al = _scalar(c.dtype, alpha)
        be = _scalar(rdt if name == "her2k" else c.dtype, beta)
        getattr(self.lib, f"{self._PREFIX}{ch}{name}")(uplo, transa, transb, m, k, _p(al), _p(a), *_estr(a), _p(b), *_estr(b),
                                                        _p(be), _p(c), *_estr(c))

    def _one(self, name, uplo, transa, alpha, a, beta, c):
        ch = CH[c.dtype]
        m = c.shape[0]
        k = a.shape[0] if (transa & TRANSPOSE) else a.shape[1]
        sdt = np.zeros(1, c.dtype).real.dtype if name == "herk" else c.dtype
        al, be = _scalar(sdt, alpha), _scalar(sdt, beta)
        getattr(self.lib, f"{self._PREFIX}{ch}{name}")(uplo, transa, m, k, _p(al), _p(a), *_estr(a), _p(be), _p(c), *_estr(c))

    def _mm(self, name, head, alpha, a, b, beta, c):
        ch = CH[c.dtype]
        m, n = c.shape
        al, be = _scalar(c.dtype, alpha), _scalar(c.dtype, beta)
        getattr(self.lib, f"{self._PREFIX}{ch}{name}")(*head, m, n, _p(al), _p(a), *_estr(a), _p(b), *_estr(b), _p(be), _p(c), *_estr(c))

    def hemm(self, side, uplo, conja, transb, alpha, a, b, beta, c): self._mm("hemm", (side, uplo, conja, transb), alpha, a, b, beta, c)
    def symm(self, side, uplo, conja, transb, alpha, a, b, beta, c): self._mm("symm", (side, uplo, conja, transb), alpha, a, b, beta, c)
    def trmm3(self, side, uplo, transa, diag, transb, alpha, a, b, beta, c): self._mm("trmm3", (side, uplo, transa, diag, transb), alpha, a, b, beta, c)

    def trmm(self, side, uplo, transa, diag, alpha, a, b):
        ch = CH[b.dtype]
        m, n = b.shape
        al = _scalar(b.dtype, alpha)
        getattr(self.lib, f"{self._PREFIX}{ch}trmm")(side, uplo, transa, diag, m, n, _p(al), _p(a), *_estr(a), _p(b), *_estr(b))

    def gemmt(self, uplo, transa, transb, alpha, a, b, beta, c): self._two("gemmt", uplo, transa, transb, alpha, a, b, beta, c)
    def syr2k(self, uplo, transa, transb, alpha, a, b, beta, c): self._two("syr2k", uplo, transa, transb, alpha, a, b, beta, c)
    def her2k(self, uplo, transa, transb, alpha, a, b, beta, c): self._two("her2k", uplo, transa, transb, alpha, a, b, beta, c)
    def syrk(self, uplo, transa, alpha, a, beta, c): self._one("syrk", uplo, transa, alpha, a, beta, c)
    def herk(self, uplo, transa, alpha, a, beta, c): self._one("herk", uplo, transa, alpha, a, beta, c)


class Oracle(_GemmtFamily):
    _PREFIX = "orc_"

    def __init__(self):
        self.lib = C.CDLL(str(build_oracle()))
        L = self.lib
        L.orc_determine_blocksize.argtypes = [ci, i64, i64, i64, i64]; L.orc_determine_blocksize.restype = i64
        L.orc_thread_range_sub.argtypes = [i64, i64, i64, i64, ci, C.POINTER(i64), C.POINTER(i64)]
        L.orc_thread_partition_2x2.argtypes = [i64, i64, i64, C.POINTER(i64), C.POINTER(i64)]
        L.orc_align_dim_to_mult.argtypes = [i64, i64]; L.orc_align_dim_to_mult.restype = i64
        L.orc_packm_panel_stride.argtypes = [i64, i64]; L.orc_packm_panel_stride.restype = i64
        L.orc_set_blksz.argtypes = [ci, i64, i64, i64, i64, i64, ci]
        for ch in "sdcz":
            getattr(L, f"orc_{ch}gemm").argtypes = [ci, ci, i64, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64]
            getattr(L, f"orc_{ch}trsm").argtypes = [ci, ci, ci, ci, i64, i64, vp, vp, i64, i64, vp, i64, i64]
            for name in ("gemmt", "syr2k", "her2k"):
                getattr(L, f"orc_{ch}{name}").argtypes = [ci, ci, ci, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64]
            for name in ("syrk", "herk"):
                getattr(L, f"orc_{ch}{name}").argtypes = [ci, ci, i64, i64, vp, vp, i64, i64, vp, vp, i64, i64]
            for name in ("hemm", "symm"):
                getattr(L, f"orc_{ch}{name}").argtypes = [ci, ci, ci, ci, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64]
            getattr(L, f"orc_{ch}trmm3").argtypes = [ci, ci, ci, ci, ci, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64]
            getattr(L, f"orc_{ch}trmm").argtypes = [ci, ci, ci, ci, i64, i64, vp, vp, i64, i64, vp, i64, i64]
            getattr(L, f"orc_{ch}packm_cxk").argtypes = [ci, i64, i64, i64, i64, vp, vp, i64, i64, vp, i64]
            getattr(L, f"orc_{ch}packm_struc_cxk").argtypes = [ci, ci, ci, ci, ci, i64, i64, i64, i64, i64, i64, vp, vp, i64, i64, vp, i64]
            getattr(L, f"orc_{ch}gemm_ukr").argtypes = [i64, i64, i64, vp, vp, vp, vp, vp, i64, i64, i64, i64]
            getattr(L, f"orc_{ch}gemmtrsm_ukr").argtypes = [ci, i64, i64, i64, vp, vp, vp, vp, vp, vp, i64, i64, i64, i64]

    def set_blksz(self, ch, mr, nr, mc, kc, nc, row_pref=0):
        self.lib.orc_set_blksz(DT[ch], mr, nr, mc, kc, nc, int(row_pref))

    def determine_blocksize(self, backward, i, dim, b_alg, b_max):
        return int(self.lib.orc_determine_blocksize(int(backward), i, dim, b_alg, b_max))

    def thread_range_sub(self, work_id, n_way, n, bf, edge_low):
        s, e = i64(), i64()
        self.lib.orc_thread_range_sub(work_id, n_way, n, bf, int(edge_low), C.byref(s), C.byref(e))
        return s.value, e.value

    def thread_partition_2x2(self, nt, w1, w2):
        a, b = i64(), i64()
        self.lib.orc_thread_partition_2x2(nt, w1, w2, C.byref(a), C.byref(b))
        return a.value, b.value

    def gemm(self, transa, transb, alpha, a, b, beta, c):
        """c := beta*c + alpha*op(a)*op(b), in place on numpy arrays with any strides."""
        ch = CH[c.dtype]
        m, n = c.shape
        k = a.shape[0] if (transa & TRANSPOSE) else a.shape[1]
        al, be = _scalar(c.dtype, alpha), _scalar(c.dtype, beta)
        getattr(self.lib, f"orc_{ch}gemm")(transa, transb, m, n, k, _p(al), _p(a), *_estr(a), _p(b), *_estr(b), _p(be), _p(c), *_estr(c))

    def trsm(self, side, uplo, transa, diag, alpha, a, b):
        ch = CH[b.dtype]
        m, n = b.shape
        al = _scalar(b.dtype, alpha)
        getattr(self.lib, f"orc_{ch}trsm")(side, uplo, transa, diag, m, n, _p(al), _p(a), *_estr(a), _p(b), *_estr(b))


class RefBlis(_GemmtFamily):
    _PREFIX = "bli_"

    """The real reference: typed API bli_?gemm / bli_?trsm (frame/3/bli_l3_tapi.c)
    plus the helper functions the oracle restates."""

    def __init__(self, threads: int | None = None):
        if not REF_SO.exists():
            raise FileNotFoundError(f"{REF_SO} missing (run oracle/build_ref.py where /root/reference exists)")
        self.lib = C.CDLL(str(REF_SO))
        L = self.lib
        L.bli_init()
        for ch in "sdcz":
            getattr(L, f"bli_{ch}gemm").argtypes = [ci, ci, i64, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64]
            getattr(L, f"bli_{ch}trsm").argtypes = [ci, ci, ci, ci, i64, i64, vp, vp, i64, i64, vp, i64, i64]
            for name in ("gemmt", "syr2k", "her2k"):
                getattr(L, f"bli_{ch}{name}").argtypes = [ci, ci, ci, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64]
            for name in ("syrk", "herk"):
                getattr(L, f"bli_{ch}{name}").argtypes = [ci, ci, i64, i64, vp, vp, i64, i64, vp, vp, i64, i64]
            for name in ("hemm", "symm"):
                getattr(L, f"bli_{ch}{name}").argtypes = [ci, ci, ci, ci, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64]
            getattr(L, f"bli_{ch}trmm3").argtypes = [ci, ci, ci, ci, ci, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64]
            getattr(L, f"bli_{ch}trmm").argtypes = [ci, ci, ci, ci, i64, i64, vp, vp, i64, i64, vp, i64, i64]
        L.bli_determine_blocksize.argtypes = [ci, i64, i64, i64, i64]; L.bli_determine_blocksize.restype = i64
        L.bli_thread_range_sub.argtypes = [i64, i64, i64, i64, C.c_bool, C.POINTER(i64), C.POINTER(i64)]
        L.bli_thread_partition_2x2.argtypes = [i64, i64, i64, C.POINTER(i64), C.POINTER(i64)]
        L.bli_thread_set_num_threads.argtypes = [i64]
        L.bli_arch_query_id.restype = ci
        L.bli_arch_string.restype = C.c_char_p; L.bli_arch_string.argtypes = [ci]
        L.bli_gks_query_cntx.restype = vp
        if threads is not None:
            L.bli_thread_set_num_threads(threads)

    def arch(self) -> str:
        return self.lib.bli_arch_string(self.lib.bli_arch_query_id()).decode()

    def set_num_threads(self, n: int):
        self.lib.bli_thread_set_num_threads(n)

    def determine_blocksize(self, backward, i, dim, b_alg, b_max):
        return int(self.lib.bli_determine_blocksize(1 if backward else 0, i, dim, b_alg, b_max))   # dir_t: BLIS_FWD=0, BLIS_BWD=1

    def thread_range_sub(self, work_id, n_way, n, bf, edge_low):
        s, e = i64(), i64()
        self.lib.bli_thread_range_sub(work_id, n_way, n, bf, bool(edge_low), C.byref(s), C.byref(e))
        return s.value, e.value

    def thread_partition_2x2(self, nt, w1, w2):
        a, b = i64(), i64()
        self.lib.bli_thread_partition_2x2(nt, w1, w2, C.byref(a), C.byref(b))
        return a.value, b.value

    def gemm(self, transa, transb, alpha, a, b, beta, c):
        ch = CH[c.dtype]
        m, n = c.shape
        k = a.shape[0] if (transa & TRANSPOSE) else a.shape[1]
        al, be = _scalar(c.dtype, alpha), _scalar(c.dtype, beta)
        getattr(self.lib, f"bli_{ch}gemm")(transa, transb, m, n, k, _p(al), _p(a), *_estr(a), _p(b), *_estr(b), _p(be), _p(c), *_estr(c))

    def trsm(self, side, uplo, transa, diag, alpha, a, b):
        ch = CH[b.dtype]
        m, n = b.shape
        al = _scalar(b.dtype, alpha)
        getattr(self.lib, f"bli_{ch}trsm")(side, uplo, transa, diag, m, n, _p(al), _p(a), *_estr(a), _p(b), *_estr(b))


SHIM_SO = ROOT / "oracle" / "_ref" / "libref_shim.so"
_MD_ARGS = [ci, ci, ci, ci, ci, ci, i64, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64]


def _gemm_md(fn, transa, transb, alpha, a, b, beta, c, comp_prec):
    """Common driver of the mixed-datatype checkers: numpy arrays of any of the four datatypes."""
    m, n = c.shape
    k = a.shape[0] if (transa & TRANSPOSE) else a.shape[1]
    if comp_prec is None:
        comp_prec = 0 if c.dtype in (np.float32, np.complex64) else 2
    al = np.array([complex(alpha).real, complex(alpha).imag], dtype=np.float64)
    be = np.array([complex(beta).real, complex(beta).imag], dtype=np.float64)
    fn(DT[CH[a.dtype]], DT[CH[b.dtype]], DT[CH[c.dtype]], comp_prec, transa, transb, m, n, k, _p(al), _p(a), *_estr(a),
       _p(b), *_estr(b), _p(be), _p(c), *_estr(c))


def oracle_gemm_md(oracle, transa, transb, alpha, a, b, beta, c, comp_prec=None):
    f = oracle.lib.orc_gemm_md
    f.argtypes = _MD_ARGS; f.restype = None
    _gemm_md(f, transa, transb, alpha, a, b, beta, c, comp_prec)


_shim = None


def ref_gemm_md(transa, transb, alpha, a, b, beta, c, comp_prec=None):
    """bli_gemm of the REAL reference on objects of different datatypes (through tests/ref_shim.c)."""
    global _shim
    if _shim is None:
        C.CDLL(str(REF_SO), mode=C.RTLD_GLOBAL)
        _shim = C.CDLL(str(SHIM_SO))
        _shim.ref_gemm_md.argtypes = _MD_ARGS; _shim.ref_gemm_md.restype = None
    _gemm_md(_shim.ref_gemm_md, transa, transb, alpha, a, b, beta, c, comp_prec)


def have_ref() -> bool:
    return REF_SO.exists()
