"""Runs in a FRESH process: loads the BLIS-side glue (which pulls in the real reference
libblis and the engine), registers the plugin, then calls the REFERENCE's own API entry points
(dgemm_, cblas_dgemm, bli_?gemm, dtrsm_, bli_?trsm) on host arrays and reports errors vs numpy
plus the number of engine kernels launched.  Printed as one JSON line."""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests"))
import gen  # noqa: E402

glue = C.CDLL(str(ROOT / "oracle" / "_ref" / "libblis_b200_glue.so"), mode=C.RTLD_GLOBAL)   # first: interposes bli_trsm_ex, bli_gemmt_ex, ...
from refblis import RefBlis, LEFT, RIGHT, LOWER, UPPER, TRANSPOSE, CONJ_TRANSPOSE, NONUNIT_DIAG, UNIT_DIAG  # noqa: E402
ref = RefBlis(threads=4)
glue.bli_plugin_register_b200.restype = C.c_int
assert glue.bli_plugin_register_b200() == -1
eng = C.CDLL(str(ROOT / "blis_b200" / "libblis_b200.so"))
eng.b200_launch_count.restype = C.c_ulonglong
L = ref.lib
out = {"arch": ref.arch()}
n0 = eng.b200_launch_count()

# --- Fortran BLAS: dgemm_ (column-major, 32-bit ints)
m, n, k = 300, 200, 150
a = gen.matrix("d", k, m, 1, "frac"); b = gen.matrix("d", k, n, 2, "frac"); c = gen.matrix("d", m, n, 3, "frac", pad=5)
want = 1.2 * c + 2.0 * (a.T @ b)
i32 = lambda v: C.byref(C.c_int(v))  # noqa: E731
f64 = lambda v: C.byref(C.c_double(v))  # noqa: E731
L.dgemm_(C.c_char_p(b"T"), C.c_char_p(b"N"), i32(m), i32(n), i32(k), f64(2.0), a.ctypes.data_as(C.c_void_p), i32(a.strides[1] // 8),
         b.ctypes.data_as(C.c_void_p), i32(b.strides[1] // 8), f64(1.2), c.ctypes.data_as(C.c_void_p), i32(c.strides[1] // 8))
out["dgemm_"] = float(np.abs(c - want).max())
out["launches_after_dgemm_"] = int(eng.b200_launch_count() - n0)

# --- CBLAS row-major: cblas_dgemm(order=101 RowMajor, NoTrans=111, Trans=112, ...)
a = gen.matrix("d", m, k, 4, "frac", "r"); b = gen.matrix("d", n, k, 5, "frac", "r"); c = gen.matrix("d", m, n, 6, "frac", "r")
want = 0.5 * c - 1.0 * (a @ b.T)
L.cblas_dgemm.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                          C.c_double, C.c_void_p, C.c_int]
L.cblas_dgemm(101, 111, 112, m, n, k, -1.0, a.ctypes.data, a.strides[0] // 8, b.ctypes.data, b.strides[0] // 8, 0.5, c.ctypes.data, c.strides[0] // 8)
out["cblas_dgemm"] = float(np.abs(c - want).max())

# --- typed BLIS API, all four datatypes (bli_?gemm)
for ch in "sdcz":
    a = gen.matrix(ch, m, k, 7, "frac"); b = gen.matrix(ch, k, n, 8, "frac", "r"); c = gen.matrix(ch, m, n, 9, "frac", "g")
    al, be = ((2.0 + 0.2j, 1.2 + 0.5j) if ch in "cz" else (2.0, 1.2))
    hi = np.complex128 if ch in "cz" else np.float64
    want = be * c.astype(hi) + al * (a.astype(hi) @ b.astype(hi))
    ref.gemm(0, 0, al, a, b, be, c)
    out[f"bli_{ch}gemm"] = float(np.abs(c - want).max())

# --- trsm through dtrsm_ and bli_?trsm (bli_trsm_ex interposed by the glue)
before = eng.b200_launch_count()
ma, nb = 257, 130
a = gen.triangular("d", ma, 10, "frac"); gen.poison_unstored(a, True); b = gen.matrix("d", ma, nb, 11, "frac")
b0 = b.copy(order="K")
L.dtrsm_(C.c_char_p(b"L"), C.c_char_p(b"L"), C.c_char_p(b"N"), C.c_char_p(b"N"), i32(ma), i32(nb), f64(2.0),
         a.ctypes.data_as(C.c_void_p), i32(a.strides[1] // 8), b.ctypes.data_as(C.c_void_p), i32(b.strides[1] // 8))
out["dtrsm_"] = float(np.abs(np.tril(np.nan_to_num(a)) @ b - 2.0 * b0).max())
for ch in "sdcz":
    a = gen.triangular(ch, nb, 12, "frac", "r"); b = gen.matrix(ch, ma, nb, 13, "frac")
    b0 = b.copy(order="K")
    al = (2.0 + 0.3j) if ch in "cz" else 2.0
    ref.trsm(RIGHT, UPPER, CONJ_TRANSPOSE if ch in "cz" else TRANSPOSE, NONUNIT_DIAG, al, a, b)
    hi = np.complex128 if ch in "cz" else np.float64
    t = np.triu(a).astype(hi).conj().T
    out[f"bli_{ch}trsm"] = float(np.abs(b.astype(hi) @ t - al * b0.astype(hi)).max())
out["launches_trsm"] = int(eng.b200_launch_count() - before)

# --- gemmt family through dsyrk_ and bli_?gemmt / herk / her2k / syr2k (bli_*_ex interposed by the glue)
before = eng.b200_launch_count()
m, k = 230, 90
a = gen.matrix("d", m, k, 20, "frac"); c = gen.matrix("d", m, m, 21, "frac"); c0 = c.copy(order="K")
L.dsyrk_(C.c_char_p(b"L"), C.c_char_p(b"N"), i32(m), i32(k), f64(2.0), a.ctypes.data_as(C.c_void_p), i32(a.strides[1] // 8),
         f64(1.2), c.ctypes.data_as(C.c_void_p), i32(c.strides[1] // 8))
want = np.where(np.tril(np.ones((m, m), bool)), 1.2 * c0 + 2.0 * (a @ a.T), c0)
out["dsyrk_"] = float(np.abs(c - want).max())
b = gen.matrix("d", k, m, 22, "frac"); c = c0.copy(order="K")
ref.gemmt(UPPER, 0, 0, 2.0, a, b, 1.2, c)
out["bli_dgemmt"] = float(np.abs(c - np.where(np.triu(np.ones((m, m), bool)), 1.2 * c0 + 2.0 * (a @ b), c0)).max())
b = gen.matrix("d", m, k, 23, "frac"); c = c0.copy(order="K")
ref.syr2k(LOWER, 0, 0, 2.0, a, b, 1.2, c)
out["bli_dsyr2k"] = float(np.abs(c - np.where(np.tril(np.ones((m, m), bool)), 1.2 * c0 + 2.0 * (a @ b.T + b @ a.T), c0)).max())
for ch in "cz":
    a = gen.matrix(ch, m, k, 24, "frac"); b = gen.matrix(ch, m, k, 25, "frac"); c0 = gen.matrix(ch, m, m, 26, "frac")
    hi = np.complex128
    low = np.tril(np.ones((m, m), bool))
    c = c0.copy(order="K"); ref.herk(LOWER, 0, 2.0, a, 1.2, c)
    w = 1.2 * c0.astype(hi) + 2.0 * (a.astype(hi) @ a.astype(hi).conj().T); w[np.diag_indices(m)] = w[np.diag_indices(m)].real
    out[f"bli_{ch}herk"] = float(np.abs(c - np.where(low, w, c0)).max())
    c = c0.copy(order="K"); al = 2.0 + 0.2j; ref.her2k(LOWER, 0, 0, al, a, b, 1.2, c)
    w = 1.2 * c0.astype(hi) + al * (a.astype(hi) @ b.astype(hi).conj().T) + np.conj(al) * (b.astype(hi) @ a.astype(hi).conj().T)
    w[np.diag_indices(m)] = w[np.diag_indices(m)].real
    out[f"bli_{ch}her2k"] = float(np.abs(c - np.where(low, w, c0)).max())
out["launches_gemmt_family"] = int(eng.b200_launch_count() - before)

# --- symm / trmm through the Fortran BLAS layer (bli_symm_ex / bli_trmm_ex interposed by the glue)
before = eng.b200_launch_count()
m, n = 210, 130
a = gen.matrix("d", m, m, 30, "frac"); b = gen.matrix("d", m, n, 31, "frac"); c = gen.matrix("d", m, n, 32, "frac"); c0 = c.copy(order="K")
asym = np.tril(a) + np.tril(a, -1).T
L.dsymm_(C.c_char_p(b"L"), C.c_char_p(b"L"), i32(m), i32(n), f64(2.0), a.ctypes.data_as(C.c_void_p), i32(a.strides[1] // 8),
         b.ctypes.data_as(C.c_void_p), i32(b.strides[1] // 8), f64(1.2), c.ctypes.data_as(C.c_void_p), i32(c.strides[1] // 8))
out["dsymm_"] = float(np.abs(c - (1.2 * c0 + 2.0 * (asym @ b))).max())
b0 = b.copy(order="K")
L.dtrmm_(C.c_char_p(b"L"), C.c_char_p(b"U"), C.c_char_p(b"T"), C.c_char_p(b"N"), i32(m), i32(n), f64(2.0),
         a.ctypes.data_as(C.c_void_p), i32(a.strides[1] // 8), b.ctypes.data_as(C.c_void_p), i32(b.strides[1] // 8))
out["dtrmm_"] = float(np.abs(b - 2.0 * (np.triu(a).T @ b0)).max())
out["launches_symm_trmm"] = int(eng.b200_launch_count() - before)

# --- dgemm_batch_ (BLAS extension, frame/compat/extra/bla_gemm_batch.c) interposed by the glue: two groups on host arrays
before = eng.b200_launch_count()
shapes = [(40, 30, 20, 3), (65, 17, 50, 2)]
As, Bs, Cs, wants = [], [], [], []
for gi, (m_, n_, k_, cnt) in enumerate(shapes):
    for j in range(cnt):
        A_ = gen.matrix("d", m_, k_, 40 + 7 * gi + j, "frac"); B_ = gen.matrix("d", k_, n_, 60 + 7 * gi + j, "frac"); C_ = gen.matrix("d", m_, n_, 80 + 7 * gi + j, "frac")
        A_, B_, C_ = (np.asfortranarray(x) for x in (A_, B_, C_))
        wants.append((1.0 + gi) * (A_ @ B_) + 0.5 * C_)
        As.append(A_); Bs.append(B_); Cs.append(C_)
ng = len(shapes)
I32 = C.c_int * ng
P = C.c_void_p * len(As)
glue.dgemm_batch_(C.c_char_p(b"NN"), C.c_char_p(b"NN"), I32(*[s_[0] for s_ in shapes]), I32(*[s_[1] for s_ in shapes]), I32(*[s_[2] for s_ in shapes]),
               (C.c_double * ng)(1.0, 2.0), P(*[x.ctypes.data for x in As]), I32(*[s_[0] for s_ in shapes]),
               P(*[x.ctypes.data for x in Bs]), I32(*[s_[2] for s_ in shapes]), (C.c_double * ng)(0.5, 0.5),
               P(*[x.ctypes.data for x in Cs]), I32(*[s_[0] for s_ in shapes]), C.byref(C.c_int(ng)), I32(*[s_[3] for s_ in shapes]))
out["dgemm_batch_"] = float(max(np.abs(c_ - w_).max() for c_, w_ in zip(Cs, wants)))
out["launches_gemm_batch"] = int(eng.b200_launch_count() - before)
out["launches_total"] = int(eng.b200_launch_count() - n0)
print(json.dumps(out))
