"""Route 2 of INTEGRATION.md built for real: the reference BLIS compiled WITH the `b200` sub-configuration.

TEST INFRASTRUCTURE (needs /root/reference; the outputs travel to the GPU box with the snapshot).

What a maintainer does by hand (docs/ConfigurationHowTo.md:615-795) is applied here by `make_overlay()` to COPIES of the
few reference files that must know about a new sub-configuration; the copies live in a scratch build directory outside
the repository (never committed, never shipped), every other source file is compiled where it lies under /root/reference:

  frame/include/bli_type_defs.h            arch_t gains BLIS_ARCH_B200                       (:1013-1019)
  frame/base/bli_arch.c                    family -> id mapping and the name "b200"          (:326-329, :358-394)
  frame/include/bli_gentconf_macro_defs.h  INSERT_GENTCONF_B200 (registers the context in bli_gks_init, frame/base/bli_gks.c:58-93,
                                           and declares bli_cntx_init_b200{,_ref} through frame/include/bli_arch_config.h:44-50)
  frame/include/bli_arch_config.h          #include "bli_family_b200.h"
  frame/3/bli_l3_oapi_ex.c                 the eleven bli_<op>_ex definitions step aside under BLIS_CONFIG_B200 -- the
                                           mechanism the sandbox uses for bli_gemm_ex (:45-49) -- so that the glue's
                                           definitions (bli_b200_glue.c, -DBLIS_B200_OVERRIDE_*) are THE bli_<op>_ex of the library
plus config/b200/{bli_cntx_init_b200.c, bli_gemm_b200_ukr.c, bli_family_b200.h, bli_kernel_defs_b200.h} from
blis_b200/blis_glue/config/b200/ and the glue itself.  Result:

  oracle/_ref/libblis_b200cfg.so           libblis whose ONLY sub-configuration is b200 (bli_arch_string -> "b200"), linked to
                                           libblis_b200.so; no LD_PRELOAD, no run-time registration
  oracle/_ref/b200cfg/test_libblis.x       the reference's own testsuite linked against it
  oracle/_ref/b200cfg/{s,d,c,z}blat3.x     the netlib level-3 BLAS testers (blastest/src, f2c'ed Fortran) linked against it,
  oracle/_ref/blastest_ref/{s,d,c,z}blat3.x  and against the plain reference (run with the plugin preloaded)
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import re
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
OUTDIR = ROOT / "oracle" / "_ref"
LIB = OUTDIR / "libblis_b200cfg.so"
BINDIR = OUTDIR / "b200cfg"
BLAT_REF_DIR = OUTDIR / "blastest_ref"
BUILD = Path(os.environ.get("B200CFG_BUILD_DIR", "/tmp/blis_b200cfg_build"))
CFG_DIR = ROOT / "blis_b200" / "blis_glue" / "config" / "b200"
GLUE_SRC = ROOT / "blis_b200" / "blis_glue" / "plugin" / "bli_b200_glue.c"
L3_EX_OPS = ("gemm", "gemmt", "her2k", "syr2k", "hemm", "symm", "trmm3", "herk", "syrk", "trmm", "trsm")

sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT))


def _sub_once(text: str, old: str, new: str, what: str) -> str:
    if text.count(old) != 1:
        raise RuntimeError(f"overlay anchor for {what} found {text.count(old)} times (reference changed?)")
    return text.replace(old, new)


def make_overlay() -> Path:
    """Patched copies of the five reference files (see module docstring) under BUILD/overlay."""
    ov = BUILD / "overlay"
    ov.mkdir(parents=True, exist_ok=True)
    # A quoted #include is resolved in the including file's own directory first, so the patched headers only win if the
    # whole of frame/include is mirrored here (scratch copies; frame/include itself is then left out of the -I list).
    for h in (REF / "frame" / "include").glob("*.h"):
        shutil.copyfile(h, ov / h.name)
    t = (REF / "frame/include/bli_type_defs.h").read_text()
    t = _sub_once(t, "\t// Generic architecture/configuration\n\tBLIS_ARCH_GENERIC,",
                  "\t// NVIDIA B200 (whole-operation level-3 engine, config/b200)\n\tBLIS_ARCH_B200,\n\n"
                  "\t// Generic architecture/configuration\n\tBLIS_ARCH_GENERIC,", "arch_t")
    (ov / "bli_type_defs.h").write_text(t)

    t = (REF / "frame/base/bli_arch.c").read_text()
    t = _sub_once(t, "\t\t// Generic microarchitecture.\n", "\t\t// NVIDIA B200.\n\t\t#ifdef BLIS_FAMILY_B200\n\t\tid = BLIS_ARCH_B200;\n\t\t#endif\n\n"
                  "\t\t// Generic microarchitecture.\n", "family -> id")
    t = _sub_once(t, '    "generic"\n};', '    "b200",\n\n    "generic"\n};', "config_name[]")
    (ov / "bli_arch.c").write_text(t)

    t = (REF / "frame/include/bli_gentconf_macro_defs.h").read_text()
    t = _sub_once(t, "// -- Generic architectures ----------------------------------------------------\n",
                  "// -- NVIDIA B200 -------------------------------------------------------------\n\n#ifdef BLIS_CONFIG_B200\n"
                  "#define INSERT_GENTCONF_B200 GENTCONF( B200, b200 )\n#else\n#define INSERT_GENTCONF_B200\n#endif\n\n"
                  "// -- Generic architectures ----------------------------------------------------\n", "INSERT_GENTCONF_B200")
    t = _sub_once(t, "INSERT_GENTCONF_SIFIVE_X280 \\\n\\\nINSERT_GENTCONF_GENERIC",
                  "INSERT_GENTCONF_SIFIVE_X280 \\\n\\\nINSERT_GENTCONF_B200 \\\n\\\nINSERT_GENTCONF_GENERIC", "INSERT_GENTCONF list")
    (ov / "bli_gentconf_macro_defs.h").write_text(t)

    t = (REF / "frame/include/bli_arch_config.h").read_text()
    t = _sub_once(t, "// -- Generic --\n\n#ifdef BLIS_FAMILY_GENERIC", "// -- NVIDIA B200 --\n\n#ifdef BLIS_FAMILY_B200\n#include \"bli_family_b200.h\"\n#endif\n\n"
                  "// -- Generic --\n\n#ifdef BLIS_FAMILY_GENERIC", "family header")
    (ov / "bli_arch_config.h").write_text(t)

    t = (REF / "frame/3/bli_l3_oapi_ex.c").read_text()
    t = _sub_once(t, "#ifdef BLIS_ENABLE_SANDBOX\nvoid PASTEMAC(gemm_def,BLIS_OAPI_EX_SUF)",
                  "#if defined(BLIS_ENABLE_SANDBOX) || defined(BLIS_CONFIG_B200)\nvoid PASTEMAC(gemm_def,BLIS_OAPI_EX_SUF)", "bli_gemm_ex guard")
    for op in L3_EX_OPS[1:]:
        head = f"\nvoid PASTEMAC({op},BLIS_OAPI_EX_SUF)\n     (\n"
        t = _sub_once(t, head, f"\n#ifdef BLIS_CONFIG_B200\nvoid PASTEMAC({op}_def,BLIS_OAPI_EX_SUF)\n#else\nvoid PASTEMAC({op},BLIS_OAPI_EX_SUF)\n#endif\n     (\n",
                      f"bli_{op}_ex guard")
    (ov / "bli_l3_oapi_ex.c").write_text(t)
    return ov


BLI_CONFIG_H = """\
/* bli_config.h of the b200 build (stands in for the file configure generates from build/bli_config.h.in when run as
   `./configure -t pthreads --enable-cblas b200`; every other value is the reference's default). */
#ifndef BLIS_CONFIG_H
#define BLIS_CONFIG_H
#define BLIS_FAMILY_B200
#define BLIS_CONFIG_B200
#define BLIS_VERSION_STRING "3.0-b200"
#define BLIS_VERSION_MAJOR 3
#define BLIS_VERSION_MINOR 0
#define BLIS_VERSION_REVISION 0
#define BLIS_ENABLE_SYSTEM
#define BLIS_ENABLE_TLS
#define BLIS_ENABLE_PTHREADS
#define BLIS_ENABLE_PTHREADS_AS_DEFAULT
#define BLIS_ENABLE_JRIR_SLAB
#define BLIS_ENABLE_PBA_POOLS
#define BLIS_ENABLE_SBA_POOLS
#define BLIS_DISABLE_MEM_TRACING
#define BLIS_DISABLE_SCALAPACK_COMPAT
#define BLIS_BLAS_INT_TYPE_SIZE 32
#define BLIS_ENABLE_BLAS
#define BLIS_ENABLE_CBLAS
#define BLIS_ENABLE_SUP_HANDLING
#define BLIS_DISABLE_MEMKIND
#define BLIS_ENABLE_TRSM_PREINVERSION
#define BLIS_ENABLE_PRAGMA_OMP_SIMD
#define BLIS_DISABLE_SANDBOX
#define BLIS_ENABLE_SHARED
#define BLIS_DISABLE_COMPLEX_RETURN_INTEL
#endif
"""


def _inc_dirs(ov: Path):
    import build_ref
    inc = BUILD / "include"
    inc.mkdir(parents=True, exist_ok=True)
    (inc / "bli_config.h").write_text(BLI_CONFIG_H)
    (inc / "bli_addon.h").write_text(build_ref.BLI_ADDON_H)
    dirs = [ov, inc, CFG_DIR, ROOT / "include"]
    for r in (REF / "frame", REF / "ref_kernels"):
        for d in [r] + sorted(x for x in r.rglob("*") if x.is_dir()):
            if build_ref.IGNORE_DIRS & set(d.relative_to(REF).parts):
                continue
            if any(d.glob("*.h")) and d != REF / "frame" / "include":
                dirs.append(d)
    return dirs


def _compile_all(ov: Path, incs):
    import build_ref
    base = ("-fPIC -std=c99 -D_POSIX_C_SOURCE=200809L -pthread -Wall -Wno-unused-function -Wfatal-errors "
            "-DBLIS_IS_BUILDING_LIBRARY -fvisibility=default")
    patched_c = {"bli_arch.c": ov / "bli_arch.c", "bli_l3_oapi_ex.c": ov / "bli_l3_oapi_ex.c"}
    jobs = []
    for s in build_ref._srcs(REF / "frame"):
        jobs.append((patched_c.get(s.name, s), "frame", f"{base} -O2"))
    cname = "-DBLIS_CNAME=b200 -DBLIS_CNAME_UPPER=B200"
    for s in sorted(CFG_DIR.glob("*.c")):
        jobs.append((s, "config_b200", f"{base} -O2 {cname}"))
    kdefs = CFG_DIR / "bli_kernel_defs_b200.h"
    for s in build_ref._srcs(REF / "ref_kernels"):
        jobs.append((s, "ref_b200", f"{base} -O2 -O3 -funsafe-math-optimizations -ffp-contract=fast -fopenmp-simd {cname} "
                                    f"-DBLIS_IN_REF_KERNEL=1 -include {kdefs}"))
    # the glue: its bli_<op>_ex definitions ARE the library's (config route); ?gemm_batch_ stays the reference's loop over bli_?gemm_ex
    jobs.append((GLUE_SRC, "glue", f"{base} -O2 -DBLIS_B200_OVERRIDE_TRSM_EX -DBLIS_B200_OVERRIDE_GEMMT_EX -DBLIS_B200_OVERRIDE_GEMM_EX"))
    objdir = BUILD / "obj"
    objdir.mkdir(parents=True, exist_ok=True)
    newest_hdr = max(p.stat().st_mtime for p in list(ov.glob("*.h")) + list(CFG_DIR.glob("*.h")) + [ROOT / "include" / "blis_b200.h"])

    def cc(job):
        src, tag, flags = job
        h = hashlib.sha1(f"{tag}:{src}".encode()).hexdigest()[:16]
        obj = objdir / f"{tag}_{src.stem}_{h}.o"
        if obj.exists() and obj.stat().st_mtime >= max(src.stat().st_mtime, newest_hdr):
            return str(obj), None
        r = subprocess.run(["gcc", *flags.split(), *incs, "-c", str(src), "-o", str(obj)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return str(obj), (None if r.returncode == 0 else f"{src}\n{r.stdout[-2000:]}")
    objs, errs = [], []
    with cf.ThreadPoolExecutor(os.cpu_count() or 4) as ex:
        for o, e in ex.map(cc, jobs):
            objs.append(o)
            if e:
                errs.append(e)
    if errs:
        raise RuntimeError("b200cfg build failed:\n" + "\n".join(errs[:4]))
    return objs


def _build_testsuite(incs) -> Path:
    exe = BINDIR / "test_libblis.x"
    objdir = BUILD / "ts_obj"
    objdir.mkdir(parents=True, exist_ok=True)
    srcs = sorted((REF / "testsuite" / "src").glob("*.c"))

    def cc(s):
        o = objdir / (s.stem + ".o")
        r = subprocess.run(["gcc", "-std=c99", "-O2", "-D_POSIX_C_SOURCE=200809L", "-pthread", *incs, f"-I{REF / 'testsuite' / 'src'}", "-c", str(s), "-o", str(o)],
                           capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(r.stderr[-2000:])
        return str(o)
    with cf.ThreadPoolExecutor(os.cpu_count() or 4) as ex:
        objs = list(ex.map(cc, srcs))
    subprocess.run(["gcc", "-pthread", *objs, "-o", str(exe), f"-L{OUTDIR}", "-lblis_b200cfg", f"-L{ROOT / 'blis_b200'}", "-lblis_b200", "-lm", "-lpthread",
                    "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath,$ORIGIN/../../../blis_b200"], check=True)
    return exe


def _build_blastest(incs, libdir: Path, libname: str, outdir: Path, extra_link=()):
    """The netlib level-3 testers (blastest/src/?blat3.c + blastest/f2c) against one library; the input files are copied
    next to the binaries (the drivers read them from stdin)."""
    outdir.mkdir(parents=True, exist_ok=True)
    objdir = BUILD / f"blat_obj_{libname}"
    objdir.mkdir(parents=True, exist_ok=True)
    flags = ["-std=c99", "-O2", "-D_POSIX_C_SOURCE=200809L", "-Wno-parentheses", "-Wno-maybe-uninitialized", f"-I{REF / 'blastest' / 'f2c'}",
             "-DHAVE_BLIS_H", *incs]

    def cc(s):
        o = objdir / (s.stem + ".o")
        if not o.exists() or o.stat().st_mtime < s.stat().st_mtime:
            r = subprocess.run(["gcc", *flags, "-c", str(s), "-o", str(o)], capture_output=True, text=True)
            if r.returncode:
                raise RuntimeError(f"{s}: {r.stderr[-1500:]}")
        return str(o)
    with cf.ThreadPoolExecutor(os.cpu_count() or 4) as ex:
        f2c = list(ex.map(cc, sorted((REF / "blastest" / "f2c").glob("*.c"))))
        drv = {ch: ex.submit(cc, REF / "blastest" / "src" / f"{ch}blat3.c") for ch in "sdcz"}
        drv = {ch: f.result() for ch, f in drv.items()}
    for ch in "sdcz":
        subprocess.run(["gcc", "-pthread", drv[ch], *f2c, "-o", str(outdir / f"{ch}blat3.x"), f"-L{libdir}", f"-l{libname}", *extra_link, "-lm", "-lpthread",
                        "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath,$ORIGIN/../../../blis_b200"], check=True)
        shutil.copyfile(REF / "blastest" / "input" / f"{ch}blat3.in", outdir / f"{ch}blat3.in")


def _derive_ukr_ops():
    """input.operations.ukr: only the level-3 MICROKERNEL modules gemm, trsm, gemmtrsm (switch 2 = "only these") --
    with the b200 context their gemm part runs in the engine-backed BLIS_GEMM_UKR slot (bli_gemm_b200_ukr.c)."""
    lines = (REF / "testsuite" / "input.operations.fast").read_text().splitlines()
    out, in_ukr = [], False
    for ln in lines:
        if ln.startswith("# --- Level-3 micro-kernels"):
            in_ukr = True
        elif ln.startswith("# --- Level-3 ---"):
            in_ukr = False
        m = re.match(r"^(\d)(\s+#\s+)(\w+)\s*$", ln)
        if in_ukr and m and m.group(3) in ("gemm", "trsm", "gemmtrsm"):
            ln = "2" + ln[1:]
        out.append(ln)
    (BINDIR / "input.operations.ukr").write_text("\n".join(out) + "\n")


def build(force: bool = False) -> Path:
    if not REF.exists():
        if LIB.exists():
            return LIB
        raise FileNotFoundError("no /root/reference and no prebuilt libblis_b200cfg.so")
    srcs = [GLUE_SRC, Path(__file__), ROOT / "include" / "blis_b200.h", *CFG_DIR.glob("*")]
    done = BINDIR / "dblat3.x"
    if LIB.exists() and done.exists() and (BLAT_REF_DIR / "dblat3.x").exists() and not force and \
            min(LIB.stat().st_mtime, done.stat().st_mtime) >= max(p.stat().st_mtime for p in srcs):
        return LIB
    from blis_b200 import build as engine_build
    engine_build.build()
    import build_ref
    build_ref.build()
    BINDIR.mkdir(parents=True, exist_ok=True)
    ov = make_overlay()
    incs = [f"-I{d}" for d in _inc_dirs(ov)]
    objs = _compile_all(ov, incs)
    rsp = BUILD / "objs.rsp"
    rsp.write_text("\n".join(objs))
    subprocess.run(["gcc", "-shared", "-o", str(LIB), f"@{rsp}", f"-L{ROOT / 'blis_b200'}", "-lblis_b200", "-lm", "-lpthread",
                    "-Wl,-rpath,$ORIGIN/../../blis_b200"], check=True)
    _build_testsuite(incs)
    _derive_ukr_ops()
    _build_blastest(incs, OUTDIR, "blis_b200cfg", BINDIR, extra_link=(f"-L{ROOT / 'blis_b200'}", "-lblis_b200"))
    # the same testers against the UNMODIFIED reference (run on the GPU box with the plugin preloaded, and as the CPU control)
    ref_incs = [f"-I{d}" for d in build_ref._inc_dirs()]
    _build_blastest(ref_incs, OUTDIR, "blis_ref", BLAT_REF_DIR)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
