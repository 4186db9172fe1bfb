#!/usr/bin/env python3
"""Compile the UNMODIFIED reference BLIS (CPU) from /root/reference into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  The result (oracle/_ref/libblis_ref.so) is the real
reference implementation of the gemm/trsm path; it pins the C restatement in
oracle/blis_oracle.c, generates the golden vectors under tests/golden/, and is
the "reference" CPU baseline that bench.py times on the GPU box's host cores.
Nothing in blis_b200/ (the product) may load it.

This is our own recipe, not the reference's build system (configure / Makefile /
common.mk are not run, no header flattening): gcc is invoked directly on the
reference's source files where they lie, with one hand-written configuration
header (oracle/_ref/include/bli_config.h, written below) and the per
sub-configuration compiler flags restated from config/<name>/make_defs.mk.

Configuration built: family "x86_64" restricted to the sub-configurations
skx, haswell, zen3, zen2, zen, generic (config_registry:11-33), pthreads
threading, BLAS + CBLAS compat layers, sup handling and trsm pre-inversion on
(the reference's defaults, build/bli_config.h.in).  A family build selects the
sub-configuration at run time from CPUID (frame/base/bli_cpuid.c:101-173), so
the library built in this container also picks the right kernels on the GPU
box's host CPU; BLIS_ARCH_TYPE / BLIS_ARCH_DEBUG (frame/base/bli_arch.c:123-190)
override / log that choice.

Outputs (git-ignored, but they travel with the gpurun snapshot):
  oracle/_ref/libblis_ref.so      the reference library
  oracle/_ref/include/*.h         generated config headers
  oracle/_ref/BUILD_INFO.json     what was built, from which tree
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import json
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
OBJ = OUT / "obj"
INC = OUT / "include"
LIB = OUT / "libblis_ref.so"
REF = Path(os.environ.get("BLIS_REFERENCE", "/root/reference"))
CC = os.environ.get("CC", "gcc")

IGNORE_DIRS = {"attic", "broken", "old", "other", "temp", "tmp", "test"}   # build/gen-make-frags/ignore_list

FAMILY = "x86_64"
CONFIGS = ["skx", "haswell", "zen3", "zen2", "zen", "generic"]
KERNEL_SETS = {"skx": "skx", "haswell": "haswell", "zen": "zen", "zen2": "zen2", "zen3": "zen3"}

# gcc flags restated from config/<name>/make_defs.mk (gcc >= 10.3 branches)
_ZEN_VEC = "-mavx2 -mfma -mfpmath=sse"
FLAGS = {
    # name: (COPTFLAGS, CKOPTFLAGS, CKVECFLAGS, CRVECFLAGS-extra)
    "x86_64": dict(copt="-O2", ckopt="-O2 -O3", ckvec="-mssse3 -mfpmath=sse -march=core2", crvec=None),
    "generic": dict(copt="-O2", ckopt="-O2 -O3", ckvec="",
                    crvec="-funsafe-math-optimizations -ffp-contract=fast"),
    "haswell": dict(copt="-O2", ckopt="-O2 -O3 -fomit-frame-pointer",
                    ckvec="-mavx2 -mfma -mfpmath=sse -march=haswell",
                    crvec="-mavx2 -mfma -mfpmath=sse -march=haswell -funsafe-math-optimizations -ffp-contract=fast"),
    "skx": dict(copt="-O2", ckopt="-O2 -O3 -fomit-frame-pointer",
                ckvec="-mavx512f -mavx512dq -mavx512bw -mavx512vl -mfpmath=sse -march=skylake-avx512",
                crvec="-march=skylake-avx512 -mno-avx512f -mno-avx512vl -mno-avx512bw -mno-avx512dq -mno-avx512cd "
                      "-funsafe-math-optimizations -ffp-contract=fast"),
    "zen": dict(copt="-O2 -fomit-frame-pointer", ckopt="-O2 -fomit-frame-pointer -O3",
                ckvec=f"{_ZEN_VEC} -march=znver1",
                crvec=f"{_ZEN_VEC} -funsafe-math-optimizations -ffp-contract=fast -march=znver1"),
    "zen2": dict(copt="-O2 -fomit-frame-pointer", ckopt="-O2 -fomit-frame-pointer -O3",
                 ckvec=f"{_ZEN_VEC} -march=znver2",
                 crvec=f"{_ZEN_VEC} -funsafe-math-optimizations -ffp-contract=fast -march=znver2"),
    "zen3": dict(copt="-O3", ckopt="-O3 -fomit-frame-pointer",
                 ckvec="-mavx2 -mfma -mfpmath=sse -march=znver3",
                 crvec="-mavx2 -mfma -funsafe-math-optimizations -ffp-contract=fast -march=znver3"),
}

BLI_CONFIG_H = """\
/* Hand-written configuration header for the oracle build of reference BLIS.
   Stands in for the file the reference's configure would generate from
   build/bli_config.h.in; every value is the reference's default. */
#ifndef BLIS_CONFIG_H
#define BLIS_CONFIG_H
#define BLIS_FAMILY_X86_64
{config_defs}
{kernel_defs}
#define BLIS_VERSION_STRING "3.0-oracle"
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

BLI_ADDON_H = """\
#ifndef BLIS_ADDON_H
#define BLIS_ADDON_H
#define BLIS_DISABLE_ADDONS
#endif
"""


def _srcs(root: Path):
    for p in sorted(root.rglob("*.c")):
        if IGNORE_DIRS & set(p.relative_to(REF).parts):
            continue
        # AMD-specific framework variants (*_amd.c) are only built with
        # ENABLE_AMD_FRAME_TWEAKS=yes (reference Makefile:244-255); vanilla build drops them.
        if root.name == "frame" and p.stem.endswith("_amd"):
            continue
        yield p


def _inc_dirs():
    dirs = [INC]
    roots = [REF / "frame"] + [REF / "config" / c for c in CONFIGS + [FAMILY]] + \
            [REF / "kernels" / k for k in KERNEL_SETS] + [REF / "ref_kernels"]
    for r in roots:
        for d in [r] + sorted(x for x in r.rglob("*") if x.is_dir()):
            if IGNORE_DIRS & set(d.relative_to(REF).parts):
                continue
            if any(d.glob("*.h")):
                dirs.append(d)
    return dirs


def _jobs():
    base = "-fPIC -std=c99 -D_POSIX_C_SOURCE=200112L -pthread -Wall -Wno-unused-function -Wfatal-errors " \
           "-DBLIS_IS_BUILDING_LIBRARY -fvisibility=default"
    jobs = []
    fam = FLAGS[FAMILY]
    for s in _srcs(REF / "frame"):
        jobs.append((s, "frame", f"{base} {fam['copt']}"))
    for c in CONFIGS:
        f = FLAGS[c]
        up = c.upper()
        cname = f"-DBLIS_CNAME={c} -DBLIS_CNAME_UPPER={up}"
        for s in _srcs(REF / "config" / c):
            jobs.append((s, f"config_{c}", f"{base} {f['copt']} {cname}"))
        kdefs = REF / "config" / c / f"bli_kernel_defs_{c}.h"
        for s in _srcs(REF / "ref_kernels"):
            jobs.append((s, f"ref_{c}", f"{base} {f['ckopt']} {f['crvec']} -fopenmp-simd {cname} "
                                        f"-DBLIS_IN_REF_KERNEL=1 -include {kdefs}"))
    for k, c in KERNEL_SETS.items():
        f = FLAGS[c]
        for s in _srcs(REF / "kernels" / k):
            jobs.append((s, f"kern_{k}", f"{base} {f['ckopt']} {f['ckvec']} -DBLIS_CNAME={c} -DBLIS_CNAME_UPPER={c.upper()}"))
    return jobs


def _compile(job, incs):
    src, tag, flags = job
    h = hashlib.sha1(f"{tag}:{src}".encode()).hexdigest()[:16]
    obj = OBJ / f"{tag}_{src.stem}_{h}.o"
    if obj.exists() and obj.stat().st_mtime >= src.stat().st_mtime:
        return obj, None
    cmd = [CC, *flags.split(), *incs, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        return obj, f"{' '.join(cmd[:6])} ... {src}\n{r.stdout[-2000:]}"
    return obj, None


def build(force: bool = False, quiet: bool = True) -> Path:
    if not REF.exists():
        if LIB.exists():
            return LIB     # prebuilt copy travelled here (GPU box): use it
        raise RuntimeError(f"{REF} not present and no prebuilt {LIB}")
    if LIB.exists() and not force:
        return LIB
    OBJ.mkdir(parents=True, exist_ok=True)
    INC.mkdir(parents=True, exist_ok=True)
    cfg = "\n".join(f"#define BLIS_CONFIG_{c.upper()}" for c in CONFIGS)
    ker = "\n".join(f"#define BLIS_KERNELS_{k.upper()}" for k in list(KERNEL_SETS) + ["generic"])
    (INC / "bli_config.h").write_text(BLI_CONFIG_H.format(config_defs=cfg, kernel_defs=ker))
    (INC / "bli_addon.h").write_text(BLI_ADDON_H)
    incs = [f"-I{d}" for d in _inc_dirs()]
    jobs = _jobs()
    objs, errs = [], []
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for obj, err in ex.map(lambda j: _compile(j, incs), jobs):
            objs.append(str(obj))
            if err:
                errs.append(err)
    if errs:
        sys.stderr.write("\n".join(errs[:5]) + f"\n... {len(errs)} file(s) failed\n")
        raise RuntimeError("reference build failed")
    rsp = OUT / "objs.rsp"
    rsp.write_text("\n".join(objs))
    r = subprocess.run([CC, "-shared", "-o", str(LIB), f"@{rsp}", "-lm", "-lpthread"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-3000:])
        raise RuntimeError("link of libblis_ref.so failed")
    (OUT / "BUILD_INFO.json").write_text(json.dumps({
        "reference": str(REF), "family": FAMILY, "configs": CONFIGS, "kernel_sets": list(KERNEL_SETS),
        "threading": "pthreads", "cc": subprocess.run([CC, "--version"], capture_output=True, text=True).stdout.splitlines()[0],
        "n_objects": len(objs)}, indent=1))
    if not quiet:
        print(f"built {LIB} from {len(objs)} objects")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, quiet=False))
