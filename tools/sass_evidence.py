"""SASS evidence for the shipped kernels: per-kernel instruction histograms from `cuobjdump -sass` of the built library
(no GPU needed).  python -m tools.sass_evidence > profiles/r01_sass_final.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "blis_b200" / "libblis_b200.so"
# mangled-name fragment -> label (the instantiation the dispatcher picks for the BASELINE shapes)
WANT = [
    ("gemm_dmma_tma_kernelILb1ELb0ELb0ELb0E", "dgemm 16384^3 default: gemm_dmma_tma_kernel<XK=1,YK=0,TRI=0,CST=0> (TMA + DMMA, 128x128x16, 6 stages)"),
    ("gemm_dmma_tma_kernelILb1ELb0ELb0ELb1E", "dgemm small k: gemm_dmma_tma_kernel<XK=1,YK=0,TRI=0,CST=1> (D staged through the TMA ring)"),
    ("gemm_ffma_tma_kernelILb1ELb0ELb0ELb0E", "sgemm default: gemm_ffma_tma_kernel<1,0,0,0> (TMA + packed FFMA2, no TF32)"),
    ("gemm_cfma_tma_kernelILb1ELb0ELb0ELb0E", "cgemm default: gemm_cfma_tma_kernel<1,0,0,0> (TMA + packed FFMA2)"),
    ("gemm_dmma_ws_kernelI7double2", "zgemm default: gemm_dmma_ws_kernel<double2,...> (warp-specialised cp.async + DMMA), first instantiation"),
    ("trsm_base_kernelIdLi64ELi64ELi256EE", "dtrsm diagonal-block solve: trsm_base_kernel<double,64,64,256>"),
]
KEY = ("DMMA", "UTMALDG", "UTMAPF", "UBLKPF", "SYNCS", "FFMA2", "FFMA", "DFMA", "LDS", "LDSM", "LDGSTS", "LDG", "STG", "HMMA", "IMMA", "UTCHMMA", "BAR", "STL", "LDL")


def main() -> int:
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    chunks = re.split(r"\n\s*Function : ", sass)
    print(f"# SASS evidence for the kernels blis_b200/libblis_b200.so ships (sm_100a, nvcc 12.9; cuobjdump -sass, {len(chunks) - 1} kernels in the library)")
    print("# per kernel: total instructions, then the counts of the mnemonics that identify the data path")
    print("# (DMMA = FP64 tensor-core MMA, UTMALDG = TMA tensor load, SYNCS = mbarrier ops, FFMA2 = packed 2xFP32 FMA;")
    print("#  HMMA/IMMA/UTCHMMA = 0 shows that no reduced-precision tensor path is used; STL/LDL = local-memory spills)")
    for frag, label in WANT:
        body = next((c for c in chunks[1:] if frag in c.split("\n", 1)[0]), None)
        if body is None:
            print(f"\n## {label}\n   NOT FOUND ({frag})")
            continue
        ops = collections.Counter()
        for ln in body.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", ln)
            if m:
                ops[m.group(1)] += 1
        total = sum(ops.values())
        print(f"\n## {label}\n   {body.splitlines()[0].strip()[:110]}")
        print(f"   instructions: {total}")
        print("   " + "  ".join(f"{k}={ops.get(k, 0)}" for k in KEY))
        print("   top: " + "  ".join(f"{k}={v}" for k, v in ops.most_common(12)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
