"""Host-side mirror of the reference's level-3 API for the gemm/trsm hot path.

Three layers, named as in BLIS so that tests read like the reference's own:

* typed API   `bli_?gemm`, `bli_?trsm`, `bli_?gemmt`, `bli_?syrk`, `bli_?herk`, `bli_?syr2k`, `bli_?her2k`
                                                 frame/3/bli_l3_tapi.c
* object API  `bli_gemm`, `bli_trsm` on `Obj`     frame/3/bli_l3_oapi.c
* BLAS compat `dgemm_`-style column-major calls   frame/compat/bla_gemm.c:127-259,
                                                 frame/compat/bla_trsm.c:126-217

All of them end in the C ABI of libblis_b200.so (include/blis_b200.h).  Operands
are torch tensors used purely as memory handles (CUDA tensors are used in place,
CPU tensors are staged by the engine) or raw integer addresses.  Error
behaviour follows the reference: invalid arguments and engine failures raise
(the C glue calls bli_abort()); there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _lib
from ._lib import (BLIS_CONJ_NO_TRANSPOSE, BLIS_CONJ_TRANSPOSE, BLIS_DCOMPLEX, BLIS_DOUBLE,
                   BLIS_FLOAT, BLIS_LEFT, BLIS_LOWER, BLIS_NO_TRANSPOSE, BLIS_NONUNIT_DIAG,
                   BLIS_RIGHT, BLIS_SCOMPLEX, BLIS_TRANSPOSE, BLIS_UNIT_DIAG, BLIS_UPPER,
                   EngineError, check)

_DT = {torch.float32: BLIS_FLOAT, torch.float64: BLIS_DOUBLE,
       torch.complex64: BLIS_SCOMPLEX, torch.complex128: BLIS_DCOMPLEX}
_CH = {"s": torch.float32, "d": torch.float64, "c": torch.complex64, "z": torch.complex128}


def _scalar_buf(dtype: torch.dtype, v):
    if dtype == torch.float32:
        return (C.c_float * 1)(float(v))
    if dtype == torch.float64:
        return (C.c_double * 1)(float(v))
    v = complex(v)
    if dtype == torch.complex64:
        return (C.c_float * 2)(v.real, v.imag)
    return (C.c_double * 2)(v.real, v.imag)


def _ptr(x) -> int:
    return x.data_ptr() if isinstance(x, torch.Tensor) else int(x)


def _bind_stream(*ts) -> None:
    """Issue the engine's work on torch's current stream of the operands' device."""
    lib = _lib.load()
    for t in ts:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            h = torch.cuda.current_stream(t.device).cuda_stream
            lib.b200_set_stream(C.c_void_p(h if h != 0 else 1))   # 1 == cudaStreamLegacy
            return
    lib.b200_set_stream(None)


# ----------------------------------------------------------------------------- typed API
def _typed_gemm(dtype):
    def f(transa, transb, m, n, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c):
        lib = _lib.load()
        _bind_stream(a, b, c)
        al, be = _scalar_buf(dtype, alpha), _scalar_buf(dtype, beta)
        rc = lib.b200_gemm(_DT[dtype], int(transa), int(transb), m, n, k, C.addressof(al),
                           _ptr(a), rs_a, cs_a, _ptr(b), rs_b, cs_b, C.addressof(be),
                           _ptr(c), rs_c, cs_c)
        check(rc, "bli_gemm")
    return f


def bli_gemm_kpanels(dtype, transa, transb, m, n, k, alpha, a_panels, rs_a, cs_a, b_panels, rs_b, cs_b, beta, c, rs_c, cs_c):
    """C := beta*C + alpha * sum_s op(A_s) op(B_s) in one launch (the pc loop of
    frame/3/gemm/bli_gemm_blk_var3.c folded into the kernel); d and z, device operands."""
    lib = _lib.load()
    _bind_stream(c)
    np_ = len(a_panels)
    assert np_ == len(b_panels)
    al, be = _scalar_buf(dtype, alpha), _scalar_buf(dtype, beta)
    pa = (C.c_void_p * np_)(*[_ptr(x) for x in a_panels])
    pb = (C.c_void_p * np_)(*[_ptr(x) for x in b_panels])
    rc = lib.b200_gemm_kpanels(_DT[dtype], int(transa), int(transb), m, n, k, np_, C.addressof(al), pa, rs_a, cs_a,
                               pb, rs_b, cs_b, C.addressof(be), _ptr(c), rs_c, cs_c)
    check(rc, "bli_gemm_kpanels")


# ----------------------------------------------------------------------------- multi-GPU (one process per GPU)
DIST_COLS, DIST_ROWS = 0, 1
DIST_AB_STATIC, DIST_TRACE = 1, 2


def partition_2x2(n_thread: int, work1: int, work2: int):
    """bli_thread_partition_2x2 through the C ABI (host arithmetic only)."""
    a, b = C.c_int64(), C.c_int64()
    _lib.load().b200_partition_2x2(n_thread, work1, work2, C.byref(a), C.byref(b))
    return a.value, b.value


def range_sub(work_id: int, n_way: int, n: int, bf: int, handle_edge_low: bool = False):
    """bli_thread_range_sub through the C ABI (host arithmetic only)."""
    a, b = C.c_int64(), C.c_int64()
    _lib.load().b200_range_sub(work_id, n_way, n, bf, int(handle_edge_low), C.byref(a), C.byref(b))
    return a.value, b.value


def dist_plan(world: int, rank: int, m: int, n: int, k: int, kb: int) -> "_lib.DistPlan":
    p = _lib.DistPlan()
    check(_lib.load().b200_dist_plan(world, rank, m, n, k, kb, C.byref(p)), "b200_dist_plan")
    return p


def dist_init(device=None) -> None:
    """Bootstrap the engine's NCCL communicator from an initialised torch.distributed process group: rank 0 draws the
    unique id (b200_dist_unique_id), torch.distributed carries the 128 bytes, every rank calls b200_dist_init."""
    import torch.distributed as dist
    lib = _lib.load()
    if device is not None:
        torch.cuda.set_device(device)
    check(lib.b200_init(-1), "b200_init")
    buf = (C.c_ubyte * 128)()
    if dist.get_rank() == 0:
        check(lib.b200_dist_unique_id(buf), "b200_dist_unique_id")
    box = [bytes(buf)]
    dist.broadcast_object_list(box, src=0)
    ident = (C.c_ubyte * 128).from_buffer_copy(box[0])
    check(lib.b200_dist_init(dist.get_world_size(), dist.get_rank(), ident), "b200_dist_init")


def dist_finalize() -> None:
    check(_lib.load().b200_dist_finalize(), "b200_dist_finalize")


def dist_gemm(dtype, m, n, k, kb, alpha, a_loc, b_loc, beta, c_loc, rs_c, cs_c, flags=0) -> None:
    lib = _lib.load()
    _bind_stream(c_loc)
    al, be = _scalar_buf(dtype, alpha), _scalar_buf(dtype, beta)
    check(lib.b200_dist_gemm(_DT[dtype], m, n, k, kb, C.addressof(al), _ptr(a_loc), _ptr(b_loc), C.addressof(be),
                             _ptr(c_loc), rs_c, cs_c, flags), "b200_dist_gemm")


def dist_register(a_loc, b_loc) -> bool:
    """Collective.  True: later dist_gemm calls on these shards with DIST_AB_STATIC pull the panels with the copy engines."""
    return _lib.load().b200_dist_register(_ptr(a_loc), _ptr(b_loc)) == _lib.BLIS_SUCCESS


def dist_unregister(a_loc, b_loc) -> None:
    _lib.load().b200_dist_unregister(_ptr(a_loc), _ptr(b_loc))


def dist_transport() -> str:
    return {0: "none", 1: "NCCL all-gather", 2: "copy-engine gets"}[_lib.load().b200_dist_transport()]


def dist_last_wait_ms() -> float:
    return float(_lib.load().b200_dist_last_wait_ms())


def dist_gemm_1d(dtype, split, root, m, n, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c_loc, rs_c, cs_c) -> None:
    lib = _lib.load()
    _bind_stream(c_loc)
    al, be = _scalar_buf(dtype, alpha), _scalar_buf(dtype, beta)
    check(lib.b200_dist_gemm_1d(_DT[dtype], split, root, m, n, k, C.addressof(al), _ptr(a), rs_a, cs_a, _ptr(b), rs_b, cs_b,
                                C.addressof(be), _ptr(c_loc), rs_c, cs_c), "b200_dist_gemm_1d")


def dist_trsm(dtype, side, uploa, transa, diaga, root, m, n, alpha, a, rs_a, cs_a, b_loc, rs_b, cs_b) -> None:
    lib = _lib.load()
    _bind_stream(b_loc)
    al = _scalar_buf(dtype, alpha)
    check(lib.b200_dist_trsm(_DT[dtype], int(side), int(uploa), int(transa), int(diaga), root, m, n, C.addressof(al),
                             _ptr(a), rs_a, cs_a, _ptr(b_loc), rs_b, cs_b), "b200_dist_trsm")


def _typed_trsm(dtype):
    def f(side, uploa, transa, diaga, m, n, alpha, a, rs_a, cs_a, b, rs_b, cs_b):
        lib = _lib.load()
        _bind_stream(a, b)
        al = _scalar_buf(dtype, alpha)
        rc = lib.b200_trsm(_DT[dtype], int(side), int(uploa), int(transa), int(diaga), m, n,
                           C.addressof(al), _ptr(a), rs_a, cs_a, _ptr(b), rs_b, cs_b)
        check(rc, "bli_trsm")
    return f


_REAL = {torch.float32: torch.float32, torch.float64: torch.float64,
         torch.complex64: torch.float32, torch.complex128: torch.float64}


def _typed_two_operand(dtype, name):
    """bli_?gemmt / bli_?syr2k / bli_?her2k (frame/3/bli_l3_tapi.c:77-112,189-220,254-296): one triangle of the
    m x m matrix C is updated; her2k's beta is real."""
    def f(uploc, transa, transb, m, k, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c):
        lib = _lib.load()
        _bind_stream(a, b, c)
        al = _scalar_buf(dtype, alpha)
        be = _scalar_buf(_REAL[dtype] if name == "her2k" else dtype, beta)
        rc = getattr(lib, "b200_" + name)(_DT[dtype], int(uploc), int(transa), int(transb), m, k, C.addressof(al),
                                          _ptr(a), rs_a, cs_a, _ptr(b), rs_b, cs_b, C.addressof(be), _ptr(c), rs_c, cs_c)
        check(rc, "bli_" + name)
    return f


def _typed_one_operand(dtype, name):
    """bli_?syrk / bli_?herk (frame/3/bli_l3_tapi.c:157-187,222-252); herk's alpha and beta are real."""
    def f(uploc, transa, m, k, alpha, a, rs_a, cs_a, beta, c, rs_c, cs_c):
        lib = _lib.load()
        _bind_stream(a, c)
        sdt = _REAL[dtype] if name == "herk" else dtype
        al, be = _scalar_buf(sdt, alpha), _scalar_buf(sdt, beta)
        rc = getattr(lib, "b200_" + name)(_DT[dtype], int(uploc), int(transa), m, k, C.addressof(al),
                                          _ptr(a), rs_a, cs_a, C.addressof(be), _ptr(c), rs_c, cs_c)
        check(rc, "bli_" + name)
    return f


def _typed_hemm(dtype, name):
    """bli_?hemm / bli_?symm (frame/3/bli_l3_tapi.c:114-155)."""
    def f(side, uploa, conja, transb, m, n, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c):
        lib = _lib.load()
        _bind_stream(a, b, c)
        al, be = _scalar_buf(dtype, alpha), _scalar_buf(dtype, beta)
        rc = getattr(lib, "b200_" + name)(_DT[dtype], int(side), int(uploa), int(conja), int(transb), m, n, C.addressof(al),
                                          _ptr(a), rs_a, cs_a, _ptr(b), rs_b, cs_b, C.addressof(be), _ptr(c), rs_c, cs_c)
        check(rc, "bli_" + name)
    return f


def _typed_trmm3(dtype):
    """bli_?trmm3 (frame/3/bli_l3_tapi.c:298-339)."""
    def f(side, uploa, transa, diaga, transb, m, n, alpha, a, rs_a, cs_a, b, rs_b, cs_b, beta, c, rs_c, cs_c):
        lib = _lib.load()
        _bind_stream(a, b, c)
        al, be = _scalar_buf(dtype, alpha), _scalar_buf(dtype, beta)
        rc = lib.b200_trmm3(_DT[dtype], int(side), int(uploa), int(transa), int(diaga), int(transb), m, n, C.addressof(al),
                            _ptr(a), rs_a, cs_a, _ptr(b), rs_b, cs_b, C.addressof(be), _ptr(c), rs_c, cs_c)
        check(rc, "bli_trmm3")
    return f


def _typed_trmm(dtype):
    """bli_?trmm (frame/3/bli_l3_tapi.c, GENTFUNC trmm): in place on B."""
    def f(side, uploa, transa, diaga, m, n, alpha, a, rs_a, cs_a, b, rs_b, cs_b):
        lib = _lib.load()
        _bind_stream(a, b)
        al = _scalar_buf(dtype, alpha)
        rc = lib.b200_trmm(_DT[dtype], int(side), int(uploa), int(transa), int(diaga), m, n,
                           C.addressof(al), _ptr(a), rs_a, cs_a, _ptr(b), rs_b, cs_b)
        check(rc, "bli_trmm")
    return f


for _ch, _dt in _CH.items():
    globals()[f"bli_{_ch}hemm"] = _typed_hemm(_dt, "hemm")
    globals()[f"bli_{_ch}symm"] = _typed_hemm(_dt, "symm")
    globals()[f"bli_{_ch}trmm3"] = _typed_trmm3(_dt)
    globals()[f"bli_{_ch}trmm"] = _typed_trmm(_dt)
    for _name in ("gemmt", "syr2k", "her2k"):
        globals()[f"bli_{_ch}{_name}"] = _typed_two_operand(_dt, _name)
    for _name in ("syrk", "herk"):
        globals()[f"bli_{_ch}{_name}"] = _typed_one_operand(_dt, _name)

bli_sgemm, bli_dgemm = _typed_gemm(torch.float32), _typed_gemm(torch.float64)
bli_cgemm, bli_zgemm = _typed_gemm(torch.complex64), _typed_gemm(torch.complex128)
bli_strsm, bli_dtrsm = _typed_trsm(torch.float32), _typed_trsm(torch.float64)
bli_ctrsm, bli_ztrsm = _typed_trsm(torch.complex64), _typed_trsm(torch.complex128)


# ----------------------------------------------------------------------------- object API
@dataclass
class Obj:
    """The part of obj_t (frame/include/bli_type_defs.h:1230-1275) this path reads:
    buffer, dims, strides and the conjtrans / uplo / diag bits of `info`."""
    buf: torch.Tensor
    conjtrans: int = BLIS_NO_TRANSPOSE
    uplo: int = 0xE0          # BLIS_DENSE
    diag: int = BLIS_NONUNIT_DIAG

    def __post_init__(self):
        if self.buf.dim() != 2:
            raise ValueError("Obj wraps a 2-D tensor")

    @property
    def dt(self): return self.buf.dtype
    @property
    def m(self): return self.buf.shape[0]
    @property
    def n(self): return self.buf.shape[1]
    @property
    def rs(self): return self.buf.stride(0)
    @property
    def cs(self): return self.buf.stride(1)

    def dims_after_trans(self):
        return (self.n, self.m) if self.conjtrans & BLIS_TRANSPOSE else (self.m, self.n)


def bli_obj_create_with_attached_buffer(t: torch.Tensor) -> Obj:
    return Obj(t)


def bli_obj_set_conjtrans(trans: int, o: Obj) -> None: o.conjtrans = trans
def bli_obj_set_uplo(uplo: int, o: Obj) -> None: o.uplo = uplo
def bli_obj_set_diag(diag: int, o: Obj) -> None: o.diag = diag


def bli_gemm(alpha, a: Obj, b: Obj, beta, c: Obj) -> None:
    """C := beta*C + alpha*trans(A)*trans(B)  (bli_gemm, frame/3/bli_l3_oapi.c).

    Checks follow bli_gemm_check (frame/3/bli_l3_check.c:37-63): conformal
    dimensions and one datatype; mixed-datatype gemm is out of scope here."""
    if not (a.dt == b.dt == c.dt):
        return bli_gemm_md(alpha, a, b, beta, c)
    ma, ka = a.dims_after_trans()
    kb, nb = b.dims_after_trans()
    if (ma, nb) != (c.m, c.n) or ka != kb:
        raise EngineError("bli_gemm: non-conformal dimensions")
    _typed_gemm(c.dt)(a.conjtrans, b.conjtrans, c.m, c.n, ka, alpha,
                      a.buf, a.rs, a.cs, b.buf, b.rs, b.cs, beta, c.buf, c.rs, c.cs)


def bli_gemm_md(alpha, a: Obj, b: Obj, beta, c: Obj, comp_prec=None) -> None:
    """Mixed-datatype gemm on objects (bli_gemm with operands of different dt, docs/MixedDatatypes.md);
    comp_prec: torch.float32 / torch.float64, default = the precision of C (bli_obj_comp_prec's default)."""
    ma, ka = a.dims_after_trans()
    kb, nb = b.dims_after_trans()
    if (ma, nb) != (c.m, c.n) or ka != kb:
        raise EngineError("bli_gemm: non-conformal dimensions")
    if comp_prec is None:
        comp_prec = _REAL[c.dt]
    lib = _lib.load()
    _bind_stream(a.buf, b.buf, c.buf)
    al, be = complex(alpha), complex(beta)
    alb, beb = (C.c_double * 2)(al.real, al.imag), (C.c_double * 2)(be.real, be.imag)
    rc = lib.b200_gemm_md(_DT[a.dt], _DT[b.dt], _DT[c.dt], 0 if comp_prec == torch.float32 else 2, int(a.conjtrans), int(b.conjtrans),
                          c.m, c.n, ka, C.addressof(alb), _ptr(a.buf), a.rs, a.cs, _ptr(b.buf), b.rs, b.cs, C.addressof(beb),
                          _ptr(c.buf), c.rs, c.cs)
    check(rc, "bli_gemm (mixed datatype)")


def gemm_batch(dtype, groups) -> None:
    """Batched gemm (?gemm_batch_, frame/compat/extra/bla_gemm_batch.c).  `groups` is a list of dicts
    {transa, transb, m, n, k, alpha, beta, a: [tensors], b: [tensors], c: [tensors]}; every tensor of a group has the
    same strides (taken from the first one).  Device problems run concurrently on the engine's stream pool."""
    import numpy as np
    lib = _lib.load()
    first = groups[0]["c"][0]
    _bind_stream(first)
    ng = len(groups)
    I32, I64 = C.c_int * ng, C.c_int64 * ng
    gs = I32(*[len(g["c"]) for g in groups])
    ta, tb = I32(*[int(g["transa"]) for g in groups]), I32(*[int(g["transb"]) for g in groups])
    mm, nn, kk = I64(*[g["m"] for g in groups]), I64(*[g["n"] for g in groups]), I64(*[g["k"] for g in groups])
    npdt = {torch.float32: np.float32, torch.float64: np.float64, torch.complex64: np.complex64, torch.complex128: np.complex128}[dtype]
    al = np.array([g["alpha"] for g in groups], dtype=npdt); be = np.array([g["beta"] for g in groups], dtype=npdt)
    def strides(key, which): return I64(*[g[key][0].stride(which) for g in groups])
    total = sum(len(g["c"]) for g in groups)
    P = C.c_void_p * total
    pa = P(*[_ptr(t) for g in groups for t in g["a"]]); pb = P(*[_ptr(t) for g in groups for t in g["b"]])
    pc = P(*[_ptr(t) for g in groups for t in g["c"]])
    rc = lib.b200_gemm_batch(_DT[dtype], ng, gs, ta, tb, mm, nn, kk, al.ctypes.data, pa, strides("a", 0), strides("a", 1),
                             pb, strides("b", 0), strides("b", 1), be.ctypes.data, pc, strides("c", 0), strides("c", 1))
    check(rc, "gemm_batch")


def bli_trsm(side: int, alpha, a: Obj, b: Obj) -> None:
    """Solve trans(A) X = alpha B (left) or X trans(A) = alpha B (right), B := X
    (bli_trsm, frame/3/bli_l3_oapi.c; checks as bli_trsm_check)."""
    if a.dt != b.dt:
        raise EngineError("bli_trsm: mixed-datatype operands are out of scope")
    if a.m != a.n:
        raise EngineError("bli_trsm: A must be square")
    if a.uplo not in (BLIS_LOWER, BLIS_UPPER):
        raise EngineError("bli_trsm: A must be triangular (uplo lower or upper)")
    if (side == BLIS_LEFT and a.m != b.m) or (side == BLIS_RIGHT and a.m != b.n):
        raise EngineError("bli_trsm: non-conformal dimensions")
    _typed_trsm(b.dt)(side, a.uplo, a.conjtrans, a.diag, b.m, b.n, alpha,
                      a.buf, a.rs, a.cs, b.buf, b.rs, b.cs)


# ----------------------------------------------------------------------------- BLAS compat
_TR = {"N": BLIS_NO_TRANSPOSE, "T": BLIS_TRANSPOSE, "C": BLIS_CONJ_TRANSPOSE}


def _blas_gemm(ch):
    dtype = _CH[ch]

    def f(transa: str, transb: str, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
        """?gemm_ (frame/compat/bla_gemm.c:127-259): column-major, rs=1, cs=ld.
        Argument errors raise like xerbla (frame/compat/check/bla_gemm_check.h)."""
        ta, tb = transa.upper(), transb.upper()
        if ta not in _TR: raise ValueError(f"{ch}gemm_: parameter 1 (transa) invalid")
        if tb not in _TR: raise ValueError(f"{ch}gemm_: parameter 2 (transb) invalid")
        if m < 0: raise ValueError(f"{ch}gemm_: parameter 3 (m) invalid")
        if n < 0: raise ValueError(f"{ch}gemm_: parameter 4 (n) invalid")
        if k < 0: raise ValueError(f"{ch}gemm_: parameter 5 (k) invalid")
        nrowa = m if ta == "N" else k
        nrowb = k if tb == "N" else n
        if lda < max(1, nrowa): raise ValueError(f"{ch}gemm_: parameter 8 (lda) invalid")
        if ldb < max(1, nrowb): raise ValueError(f"{ch}gemm_: parameter 10 (ldb) invalid")
        if ldc < max(1, m): raise ValueError(f"{ch}gemm_: parameter 13 (ldc) invalid")
        _typed_gemm(dtype)(_TR[ta], _TR[tb], m, n, k, alpha, a, 1, lda, b, 1, ldb, beta, c, 1, ldc)
    return f


def _blas_trsm(ch):
    dtype = _CH[ch]

    def f(side: str, uplo: str, transa: str, diag: str, m, n, alpha, a, lda, b, ldb):
        """?trsm_ (frame/compat/bla_trsm.c:126-217)."""
        s, u, t, d = side.upper(), uplo.upper(), transa.upper(), diag.upper()
        if s not in "LR": raise ValueError(f"{ch}trsm_: parameter 1 (side) invalid")
        if u not in "LU": raise ValueError(f"{ch}trsm_: parameter 2 (uplo) invalid")
        if t not in _TR: raise ValueError(f"{ch}trsm_: parameter 3 (transa) invalid")
        if d not in "NU": raise ValueError(f"{ch}trsm_: parameter 4 (diag) invalid")
        if m < 0: raise ValueError(f"{ch}trsm_: parameter 5 (m) invalid")
        if n < 0: raise ValueError(f"{ch}trsm_: parameter 6 (n) invalid")
        nrowa = m if s == "L" else n
        if lda < max(1, nrowa): raise ValueError(f"{ch}trsm_: parameter 9 (lda) invalid")
        if ldb < max(1, m): raise ValueError(f"{ch}trsm_: parameter 11 (ldb) invalid")
        _typed_trsm(dtype)(BLIS_LEFT if s == "L" else BLIS_RIGHT,
                           BLIS_LOWER if u == "L" else BLIS_UPPER, _TR[t],
                           BLIS_UNIT_DIAG if d == "U" else BLIS_NONUNIT_DIAG,
                           m, n, alpha, a, 1, lda, b, 1, ldb)
    return f


sgemm_, dgemm_, cgemm_, zgemm_ = (_blas_gemm(ch) for ch in "sdcz")
strsm_, dtrsm_, ctrsm_, ztrsm_ = (_blas_trsm(ch) for ch in "sdcz")


# ----------------------------------------------------------------------------- info / tuning
def info() -> str:
    return _lib.load().b200_info().decode()


def blksz(dtype: torch.dtype, which: str) -> int:
    idx = {"MR": 0, "NR": 1, "MC": 2, "KC": 3, "NC": 4}[which]
    return int(_lib.load().b200_blksz(_DT[dtype], idx))


def measure_peak(kind: str, millis: int = 200) -> float:
    """TFLOP/s of the DFMA / DMMA / FFMA pipe microbenchmark."""
    k = {"dfma": 0, "dmma": 1, "ffma": 2, "ffma2": 3}[kind]
    v = float(_lib.load().b200_measure_peak(k, millis))
    if v < 0:
        raise EngineError("peak microbenchmark failed (no GPU?)")
    return v


def set_option(key: str, value: int) -> None:
    check(_lib.load().b200_set_option(key.encode(), int(value)), f"set_option({key})")


def launch_count() -> int:
    """Kernels launched by the engine so far (for bench.py's gpu_launches)."""
    return int(_lib.load().b200_launch_count())


def last_kernel() -> str:
    """Name of the last kernel this thread launched through the engine."""
    return _lib.load().b200_last_kernel().decode()


def kernel_stats(reset: bool = False) -> dict:
    """{kernel name: launches} since the last reset."""
    lib = _lib.load()
    n = int(lib.b200_kernel_stats(None, 0, 0))
    buf = C.create_string_buffer(n + 1)
    lib.b200_kernel_stats(buf, n + 1, 1 if reset else 0)
    out = {}
    for ln in buf.value.decode().splitlines():
        k, _, v = ln.rpartition("\t")
        if k:
            out[k] = int(v)
    return out


def sync() -> None:
    check(_lib.load().b200_sync(), "b200_sync")
