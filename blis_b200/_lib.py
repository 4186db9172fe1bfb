"""ctypes loader for libblis_b200.so (the C-ABI engine).

There is deliberately no fallback: if the CUDA library is missing or does not
load, importing the compute API raises.  `oracle/` is never imported from here.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libblis_b200.so"

# BLIS enum values (frame/include/bli_type_defs.h:278-455), same as include/blis_b200.h
BLIS_FLOAT, BLIS_SCOMPLEX, BLIS_DOUBLE, BLIS_DCOMPLEX = 0, 1, 2, 3
BLIS_NO_TRANSPOSE, BLIS_TRANSPOSE, BLIS_CONJ_NO_TRANSPOSE, BLIS_CONJ_TRANSPOSE = 0x00, 0x08, 0x10, 0x18
BLIS_UPPER, BLIS_LOWER = 0x60, 0xC0
BLIS_LEFT, BLIS_RIGHT = 0, 1
BLIS_NONUNIT_DIAG, BLIS_UNIT_DIAG = 0x000, 0x100
BLIS_SUCCESS, BLIS_FAILURE = -1, -2

# every symbol include/blis_b200.h declares
EXPORTS = [
    "b200_init", "b200_finalize", "b200_last_error", "b200_device_count", "b200_info",
    "b200_set_stream", "b200_get_stream", "b200_sync", "b200_malloc_pinned", "b200_free_pinned", "b200_pointer_kind",
    "b200_gemm", "b200_gemm_kpanels", "b200_sgemm", "b200_dgemm", "b200_cgemm", "b200_zgemm",
    "b200_trsm", "b200_strsm", "b200_dtrsm", "b200_ctrsm", "b200_ztrsm",
    "b200_gemmt", "b200_syrk", "b200_herk", "b200_syr2k", "b200_her2k",
    "b200_hemm", "b200_symm", "b200_trmm3", "b200_trmm", "b200_gemm_md", "b200_gemm_batch",
    "b200_partition_2x2", "b200_range_sub", "b200_dist_plan", "b200_dist_unique_id", "b200_dist_init", "b200_dist_finalize",
    "b200_dist_gemm", "b200_dist_last_wait_ms", "b200_dist_register", "b200_dist_unregister", "b200_dist_transport", "b200_dist_gemm_1d", "b200_dist_trsm",
    "b200_blksz", "b200_measure_peak", "b200_launch_count", "b200_set_option", "b200_last_kernel", "b200_kernel_stats",
    "b200_splitk_plan", "b200_trsm_upload_plan", "b200_trsm_rowblock_plan",
]

class DistPlan(C.Structure):
    """b200_dist_plan_t (include/blis_b200.h)."""
    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("pr", C.c_int), ("pc", C.c_int), ("i", C.c_int), ("j", C.c_int),
                ("m0", C.c_int64), ("m1", C.c_int64), ("n0", C.c_int64), ("n1", C.c_int64), ("kb", C.c_int64),
                ("L", C.c_int), ("T", C.c_int), ("steps", C.c_int), ("na", C.c_int), ("nb", C.c_int)]


_lib = None


class EngineError(RuntimeError):
    """Raised when the engine returns BLIS_FAILURE (the BLIS glue aborts instead)."""


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m blis_b200.build` "
            "(nvcc, sm_100a). blis_b200 has no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    i64, vp, ci = C.c_int64, C.c_void_p, C.c_int
    lib.b200_init.argtypes = [ci]; lib.b200_init.restype = ci
    lib.b200_finalize.argtypes = []; lib.b200_finalize.restype = None
    lib.b200_last_error.argtypes = []; lib.b200_last_error.restype = C.c_char_p
    lib.b200_device_count.argtypes = []; lib.b200_device_count.restype = ci
    lib.b200_info.argtypes = []; lib.b200_info.restype = C.c_char_p
    lib.b200_set_stream.argtypes = [vp]; lib.b200_set_stream.restype = None
    lib.b200_get_stream.argtypes = []; lib.b200_get_stream.restype = vp
    lib.b200_sync.argtypes = []; lib.b200_sync.restype = ci
    gemm_tail = [i64, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64]
    lib.b200_gemm.argtypes = [ci, ci, ci] + gemm_tail; lib.b200_gemm.restype = ci
    for ch in "sdcz":
        f = getattr(lib, f"b200_{ch}gemm"); f.argtypes = [ci, ci] + gemm_tail; f.restype = ci
    lib.b200_gemm_kpanels.argtypes = [ci, ci, ci, i64, i64, i64, ci, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64]
    lib.b200_gemm_kpanels.restype = ci
    trsm_tail = [i64, i64, vp, vp, i64, i64, vp, i64, i64]
    lib.b200_trsm.argtypes = [ci, ci, ci, ci, ci] + trsm_tail; lib.b200_trsm.restype = ci
    for ch in "sdcz":
        f = getattr(lib, f"b200_{ch}trsm"); f.argtypes = [ci, ci, ci, ci] + trsm_tail; f.restype = ci
    two_op = [ci, ci, ci, ci, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64]     # dt uplo transa transb m k ...
    one_op = [ci, ci, ci, i64, i64, vp, vp, i64, i64, vp, vp, i64, i64]                        # dt uplo transa m k ...
    for name in ("gemmt", "syr2k", "her2k"):
        f = getattr(lib, f"b200_{name}"); f.argtypes = two_op; f.restype = ci
    for name in ("syrk", "herk"):
        f = getattr(lib, f"b200_{name}"); f.argtypes = one_op; f.restype = ci
    mm_tail = [i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64]
    lib.b200_hemm.argtypes = [ci, ci, ci, ci, ci] + mm_tail; lib.b200_hemm.restype = ci          # dt side uplo conja transb
    lib.b200_symm.argtypes = [ci, ci, ci, ci, ci] + mm_tail; lib.b200_symm.restype = ci
    lib.b200_trmm3.argtypes = [ci, ci, ci, ci, ci, ci] + mm_tail; lib.b200_trmm3.restype = ci    # dt side uplo transa diag transb
    lib.b200_trmm.argtypes = [ci, ci, ci, ci, ci] + trsm_tail; lib.b200_trmm.restype = ci        # dt side uplo transa diag
    lib.b200_gemm_md.argtypes = [ci, ci, ci, ci, ci, ci, i64, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64]
    lib.b200_gemm_md.restype = ci
    lib.b200_gemm_batch.argtypes = [ci, ci] + [vp] * 17; lib.b200_gemm_batch.restype = ci
    pi64 = C.POINTER(i64)
    lib.b200_partition_2x2.argtypes = [i64, i64, i64, pi64, pi64]; lib.b200_partition_2x2.restype = None
    lib.b200_range_sub.argtypes = [i64, i64, i64, i64, ci, pi64, pi64]; lib.b200_range_sub.restype = None
    lib.b200_dist_plan.argtypes = [ci, ci, i64, i64, i64, i64, C.POINTER(DistPlan)]; lib.b200_dist_plan.restype = ci
    lib.b200_dist_unique_id.argtypes = [vp]; lib.b200_dist_unique_id.restype = ci
    lib.b200_dist_init.argtypes = [ci, ci, vp]; lib.b200_dist_init.restype = ci
    lib.b200_dist_finalize.argtypes = []; lib.b200_dist_finalize.restype = ci
    lib.b200_dist_gemm.argtypes = [ci, i64, i64, i64, i64, vp, vp, vp, vp, vp, i64, i64, ci]; lib.b200_dist_gemm.restype = ci
    lib.b200_dist_last_wait_ms.argtypes = []; lib.b200_dist_last_wait_ms.restype = C.c_double
    lib.b200_dist_register.argtypes = [vp, vp]; lib.b200_dist_register.restype = ci
    lib.b200_dist_unregister.argtypes = [vp, vp]; lib.b200_dist_unregister.restype = ci
    lib.b200_dist_transport.argtypes = []; lib.b200_dist_transport.restype = ci
    lib.b200_dist_gemm_1d.argtypes = [ci, ci, ci, i64, i64, i64, vp, vp, i64, i64, vp, i64, i64, vp, vp, i64, i64]
    lib.b200_dist_gemm_1d.restype = ci
    lib.b200_dist_trsm.argtypes = [ci, ci, ci, ci, ci, ci, i64, i64, vp, vp, i64, i64, vp, i64, i64]; lib.b200_dist_trsm.restype = ci
    lib.b200_blksz.argtypes = [ci, ci]; lib.b200_blksz.restype = i64
    lib.b200_measure_peak.argtypes = [ci, ci]; lib.b200_measure_peak.restype = C.c_double
    lib.b200_launch_count.argtypes = []; lib.b200_launch_count.restype = C.c_ulonglong
    lib.b200_set_option.argtypes = [C.c_char_p, C.c_longlong]; lib.b200_set_option.restype = ci
    lib.b200_last_kernel.argtypes = []; lib.b200_last_kernel.restype = C.c_char_p
    lib.b200_kernel_stats.argtypes = [C.c_char_p, C.c_size_t, ci]; lib.b200_kernel_stats.restype = C.c_size_t
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != BLIS_SUCCESS:
        msg = load().b200_last_error().decode(errors="replace")
        raise EngineError(f"{what} failed: {msg}")
