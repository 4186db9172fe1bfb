"""Build the reference's OWN testsuite binary (testsuite/src/*.c, unmodified) against the
reference library built by oracle/build_ref.py, and derive its input files.

TEST INFRASTRUCTURE.  Outputs go to oracle/_ref/testsuite/ (git-ignored, travels to the GPU box):
  test_libblis.x            the reference testsuite driver, linked to libblis_ref.so only
  input.general.<tag>       derived from testsuite/input.general.fast (sizes / datatypes edited)
  input.operations.l3       derived from testsuite/input.operations.fast: only gemm, trsm and the gemmt family enabled
                            (switch value 2), all transa/transb and side/uplo/trans/diag combinations
On the GPU box the binary is run twice by tests/test_blis_dropin_gpu.py: as is (CPU reference) and with
LD_PRELOAD=libblis_b200_glue.so BLIS_B200_PLUGIN=1 (gemm/trsm served by the B200 engine).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
OUT = ROOT / "oracle" / "_ref" / "testsuite"
EXE = OUT / "test_libblis.x"
L3_OPS = ("gemm", "trsm", "gemmt", "syrk", "herk", "syr2k", "her2k", "hemm", "symm", "trmm", "trmm3")      # the operations the engine serves


def _derive_inputs():
    gen = (REF / "testsuite" / "input.general.fast").read_text()
    ops = (REF / "testsuite" / "input.operations.fast").read_text().splitlines()
    # only the served operations of the "Level-3" section, every parameter combination
    out, in_l3, i = [], False, 0
    while i < len(ops):
        ln = ops[i]
        if ln.startswith("# --- Level-3 ---"):
            in_l3 = True
        m = re.match(r"^(\d)(\s+#\s+)(\w+)\s*$", ln)
        if in_l3 and m and m.group(3) in L3_OPS:
            out.append("2" + ln[1:])
            i += 1
            while i < len(ops) and ops[i].strip() and not re.match(r"^\d\s+#\s+\w+\s*$", ops[i]):
                if "parameters:" in ops[i]:
                    npar = len(ops[i].split("#")[0].strip())
                    out.append("?" * npar + ops[i][npar:])
                else:
                    out.append(ops[i])
                i += 1
            continue
        out.append(ln)
        i += 1
    (OUT / "input.operations.l3").write_text("\n".join(out) + "\n")

    def general(tag, size, dts, mixed=False, app_threads=1):
        g = gen
        if app_threads != 1:
            # "Simulate application-level threading" (testsuite/src/test_libblis.c: n testsuite threads share the
            # experiments; every thread calls the BLIS API concurrently)
            g = re.sub(r"^1(\s+# Simulate application-level threading:)", f"{app_threads}\\1", g, flags=re.M)
        if mixed:
            g = re.sub(r"^0(\s+# Test gemm with mixed-domain operands\?)", "1\\1", g, flags=re.M)
            g = re.sub(r"^0(\s+# Test gemm with mixed-precision operands\?)", "1\\1", g, flags=re.M)
        g = re.sub(r"^\d+(\s+# Problem size: first to test)", f"{size}\\1", g, flags=re.M)
        g = re.sub(r"^\d+(\s+# Problem size: maximum to test)", f"{size}\\1", g, flags=re.M)
        g = re.sub(r"^\d+(\s+# Problem size: increment between experiments)", f"{size}\\1", g, flags=re.M)
        g = re.sub(r"^\w+(\s+# Datatype\(s\) to test:)", f"{dts}\\1", g, flags=re.M)
        g = re.sub(r"^1(\s+#\s+1m\s)", "0\\1", g, flags=re.M)          # native execution only (1m is out of scope)
        (OUT / f"input.general.{tag}").write_text(g)
    general("n100", 100, "sdcz")
    general("n1000d", 1000, "d")
    general("n100mixed", 100, "sdcz", mixed=True)
    general("n100t4", 100, "sdcz", app_threads=4)
    # gemm only, for the mixed-datatype run (the other operations have no mixed-datatype variants)
    l3 = (OUT / "input.operations.l3").read_text().splitlines()
    keep, in_l3 = [], False
    for ln in l3:
        if ln.startswith("# --- Level-3 ---"):
            in_l3 = True
        m = re.match(r"^(\d)(\s+#\s+)(\w+)\s*$", ln)
        if in_l3 and m and m.group(3) != "gemm":
            ln = "0" + ln[1:]
        keep.append(ln)
    (OUT / "input.operations.gemm").write_text("\n".join(keep) + "\n")


def build(force: bool = False) -> Path:
    if not REF.exists():
        if EXE.exists():
            return EXE
        raise FileNotFoundError("no /root/reference and no prebuilt testsuite binary")
    if EXE.exists() and not force:
        OUT.mkdir(parents=True, exist_ok=True)
        _derive_inputs()
        return EXE
    sys.path.insert(0, str(ROOT / "oracle"))
    import build_ref
    lib = build_ref.build()
    OUT.mkdir(parents=True, exist_ok=True)
    (OUT / "obj").mkdir(exist_ok=True)
    incs = [f"-I{d}" for d in build_ref._inc_dirs()] + [f"-I{REF / 'testsuite' / 'src'}"]
    srcs = sorted((REF / "testsuite" / "src").glob("*.c"))

    def cc(s):
        o = OUT / "obj" / (s.stem + ".o")
        r = subprocess.run(["gcc", "-std=c99", "-O2", "-D_POSIX_C_SOURCE=200112L", "-pthread", *incs, "-c", str(s), "-o", str(o)],
                           capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(r.stderr[-2000:])
        return str(o)
    with cf.ThreadPoolExecutor(os.cpu_count() or 4) as ex:
        objs = list(ex.map(cc, srcs))
    subprocess.run(["gcc", "-pthread", *objs, "-o", str(EXE), f"-L{lib.parent}", "-lblis_ref", "-lm", "-lpthread",
                    "-Wl,-rpath,$ORIGIN/.."], check=True)
    _derive_inputs()
    return EXE


if __name__ == "__main__":
    print(build(force=True))
