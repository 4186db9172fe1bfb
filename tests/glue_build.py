"""Build the BLIS-side glue (blis_b200/blis_glue/plugin/bli_b200_glue.c) against the
reference's headers.  Only possible where /root/reference is mounted; the result
(oracle/_ref/libblis_b200_glue.so) travels with the snapshot to the GPU box.

TEST INFRASTRUCTURE: in a real installation the same file is compiled by the
plugin / config/b200 build against the installed blis.h."""
from __future__ import annotations

import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
GLUE_SRC = ROOT / "blis_b200" / "blis_glue" / "plugin" / "bli_b200_glue.c"
CFG_SRC = ROOT / "blis_b200" / "blis_glue" / "config" / "b200" / "bli_cntx_init_b200.c"
GLUE_SO = ROOT / "oracle" / "_ref" / "libblis_b200_glue.so"


# Second flavour of the glue: NOTHING interposed.  The only binding is the one BLIS itself offers a plugin -- the
# whole-operation gemm handler installed in the gemmsup_oft slot of the active context (frame/3/bli_l3_sup_oft.h:46-59,
# bli_gemm_ex -> bli_gemmsup, frame/3/bli_l3_oapi_ex.c:76-77) -- so every dgemm_/cblas_dgemm/bli_?gemm call of an
# unmodified libblis reaches the engine through bli_gemmsup_b200, and every other operation stays on the CPU.
GLUE_SUP_SO = ROOT / "oracle" / "_ref" / "libblis_b200_glue_sup.so"


def build(force: bool = False) -> Path:
    if not Path("/root/reference").exists():
        if GLUE_SO.exists():
            return GLUE_SO
        raise FileNotFoundError("no /root/reference and no prebuilt glue library")
    if GLUE_SO.exists() and GLUE_SUP_SO.exists() and not force and min(GLUE_SO.stat().st_mtime, GLUE_SUP_SO.stat().st_mtime) >= GLUE_SRC.stat().st_mtime:
        return GLUE_SO
    sys.path.insert(0, str(ROOT / "oracle"))
    import build_ref
    build_ref.build()
    incs = [f"-I{d}" for d in build_ref._inc_dirs()] + [f"-I{ROOT / 'include'}"]
    for so, defs in ((GLUE_SO, ["-DBLIS_B200_OVERRIDE_TRSM_EX", "-DBLIS_B200_OVERRIDE_GEMMT_EX", "-DBLIS_B200_OVERRIDE_GEMM_EX", "-DBLIS_B200_OVERRIDE_GEMM_BATCH"]),
                     (GLUE_SUP_SO, [])):
        cmd = ["gcc", "-std=c99", "-O2", "-fPIC", "-shared", "-D_POSIX_C_SOURCE=200809L", "-Wall", "-Wno-unused-function",
               *defs, *incs, str(GLUE_SRC), "-o", str(so),
               f"-L{ROOT / 'blis_b200'}", "-lblis_b200", f"-L{so.parent}", "-lblis_ref",
               "-Wl,-rpath,$ORIGIN/../../blis_b200", "-Wl,-rpath,$ORIGIN"]
        subprocess.run(cmd, check=True)
    return GLUE_SO


SHIM_SRC = ROOT / "tests" / "ref_shim.c"
SHIM_SO = ROOT / "oracle" / "_ref" / "libref_shim.so"


def build_shim(force: bool = False) -> Path:
    """tests/ref_shim.c -> oracle/_ref/libref_shim.so (object-API calls into the real reference)."""
    if not Path("/root/reference").exists():
        if SHIM_SO.exists():
            return SHIM_SO
        raise FileNotFoundError("no /root/reference and no prebuilt reference shim")
    if SHIM_SO.exists() and not force and SHIM_SO.stat().st_mtime >= SHIM_SRC.stat().st_mtime:
        return SHIM_SO
    sys.path.insert(0, str(ROOT / "oracle"))
    import build_ref
    build_ref.build()
    incs = [f"-I{d}" for d in build_ref._inc_dirs()]
    cmd = ["gcc", "-std=c99", "-O2", "-fPIC", "-shared", "-D_POSIX_C_SOURCE=200809L", "-Wall", "-Wno-unused-function", *incs,
           str(SHIM_SRC), "-o", str(SHIM_SO), f"-L{SHIM_SO.parent}", "-lblis_ref", "-Wl,-rpath,$ORIGIN"]
    subprocess.run(cmd, check=True)
    return SHIM_SO


def syntax_check_config() -> None:
    """config/b200/bli_cntx_init_b200.c must compile against the reference headers (with the
    b200 names a maintainer adds to bli_arch_config.h declared here)."""
    sys.path.insert(0, str(ROOT / "oracle"))
    import build_ref
    incs = [f"-I{d}" for d in build_ref._inc_dirs()] + [f"-I{ROOT / 'include'}"]
    cmd = ["gcc", "-std=c99", "-fsyntax-only", "-D_POSIX_C_SOURCE=200112L", "-Wall", "-Wno-unused-function",
           "-include", str(ROOT / "tests" / "b200_arch_decls.h"), *incs, str(CFG_SRC)]
    subprocess.run(cmd, check=True)


if __name__ == "__main__":
    print(build(force=True))
    print(build_shim(force=True))
