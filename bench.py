#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 level-3 engine (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--op both|dgemm|dtrsm]

The metric (BASELINE.json) has two halves, and the ONE JSON line carries both:

* dgemm (G1, BASELINE configs[1]): m=n=k=16384 column-major fp64, alpha=2.0 beta=1.2
  (testsuite/src/test_gemm.c:213-214), inputs uniform[-1,1] normalised as libblis_test_mobj_randomize does.
  This is the line's `value` / `ms_per_step` / `roofline` / `e2e`.
* dtrsm (T1, BASELINE configs[3]): left/lower/notrans/nonunit m=32768 n=8192, alpha=2.0 -- the `dtrsm` object
  (value, ms_per_step, roofline, kernels, e2e), strong-scaled over column blocks of B at N > 1.

N > 1 (torchrun): `value` is WEAK scaling -- every rank owns a 16384 x 16384 block of C of a larger product, 2D block
decomposition with NCCL all-gather of the A/B k-panels (blis_b200/dist.py); the `strong` object is the literal reading of
the metric (ONE 16384^3 product over all N GPUs); `check` is the post-timing parity check of the distributed result
(bit-for-bit against the same k-panel schedule replayed by the single-GPU engine on independently gathered panels, plus
the testsuite residual).

`roofline.kernel(s)` are read back from the engine (b200_kernel_stats): the kernels that actually ran in the timed region.
`e2e` at N = 1 is the REFERENCE's own Fortran entry point `dgemm_` (oracle/_ref/libblis_ref.so, the unmodified reference
library, used here only as the API front) with the B200 plugin registered (blis_glue), on pinned HOST operands: H2D of
A, B, C and D2H of C inside the timed region.  `cpu_baseline` is the reference BLIS on the host cores, run in a separate
process (`--impl reference --cpu-sample`) so that the plugin registered in this process cannot serve it.
`--impl reference` times only the CPU reference, at the full 16384^3 whenever the predicted time fits the budget.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "dgemm/dtrsm GFLOPS at n=16384 (1/2/4/8 B200) and % of FP64 peak vs BLIS host CPU"
N_DEFAULT = 16384
ALPHA, BETA = 2.0, 1.2
GLUE_SO = ROOT / "oracle" / "_ref" / "libblis_b200_glue.so"
REF_SO = ROOT / "oracle" / "_ref" / "libblis_ref.so"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- reference (CPU) arm
def _host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def _ref_lib():
    sys.path.insert(0, str(ROOT / "tests"))
    from refblis import RefBlis, have_ref
    if not have_ref():
        sys.path.insert(0, str(ROOT / "oracle"))
        import build_ref
        build_ref.build()
    return RefBlis()


def _pick_ref_subconfig() -> str:
    """The reference picks its sub-configuration with bli_cpuid_query_id (frame/base/bli_cpuid.c).  On the GPU boxes (KVM
    guests whose CPU brand string has no model number) its FMA-unit probe (vpu_count, bli_cpuid.c:930) cannot tell and it
    settles for 'haswell' on an AVX-512 host.  To give the reference its best kernels, both the auto-detected
    sub-configuration and BLIS_ARCH_TYPE=skx (the reference's own override, frame/base/bli_arch.c:128-196) are probed in
    throw-away processes on a small dgemm and the faster one is used.  Returns a note for the JSON line."""
    global _SUBCONFIG_NOTE
    if _SUBCONFIG_NOTE is not None:
        return _SUBCONFIG_NOTE
    _SUBCONFIG_NOTE = _pick_ref_subconfig_once()
    return _SUBCONFIG_NOTE


_SUBCONFIG_NOTE = None


def _pick_ref_subconfig_once() -> str:
    if "BLIS_ARCH_TYPE" in os.environ:
        return f"BLIS_ARCH_TYPE={os.environ['BLIS_ARCH_TYPE']} (set by the caller)"
    try:
        flags = Path("/proc/cpuinfo").read_text()
    except OSError:
        return "auto-detected"
    if "avx512f" not in flags or "avx512dq" not in flags or "avx512bw" not in flags or "avx512vl" not in flags:
        return "auto-detected (no AVX-512 on this host)"
    res = {}
    for arch in ("", "skx"):
        env = dict(os.environ)
        if arch:
            env["BLIS_ARCH_TYPE"] = arch
        try:
            r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--arch-probe"], env=env,
                               capture_output=True, text=True, timeout=120)
            res[arch] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 and r.stdout.strip() else None
        except (subprocess.TimeoutExpired, ValueError, OSError):
            res[arch] = None
    auto, skx = res.get(""), res.get("skx")
    if skx and (not auto or skx["gflops"] > 1.03 * auto["gflops"]):
        os.environ["BLIS_ARCH_TYPE"] = "skx"
        return (f"BLIS_ARCH_TYPE=skx override ({skx['gflops']:.0f} GFLOPS at 3072^3) because the reference's auto-detection picks "
                f"'{auto['arch'] if auto else '?'}' on this VM ({auto['gflops']:.0f} GFLOPS)" if auto else "BLIS_ARCH_TYPE=skx override")
    return f"auto-detected '{auto['arch']}'" + (f" (skx override probed: {skx['gflops']:.0f} vs {auto['gflops']:.0f} GFLOPS)" if skx else "") if auto else "auto-detected"


def _arch_probe() -> int:
    import numpy as np
    ref = _ref_lib()
    ref.set_num_threads(_host_cores())
    n = 3072
    rng = np.random.default_rng(1)
    a, b, c = (np.asfortranarray(rng.uniform(-1, 1, (n, n))) for _ in range(3))
    ref.gemm(0, 0, ALPHA, a, b, BETA, c)
    t0 = time.perf_counter(); ref.gemm(0, 0, ALPHA, a, b, BETA, c); dt = time.perf_counter() - t0
    print(json.dumps({"arch": ref.arch(), "gflops": 2.0 * n ** 3 / dt / 1e9}))
    return 0


def cpu_reference(op: str, budget_s: float, max_steps: int, n_full: int = N_DEFAULT):
    """Reference BLIS (oracle/_ref, every host core) on the workload: the full size when `max_steps` steps are predicted to
    fit `budget_s`, else the largest power-of-two fraction that does.  Returns a dict(value, ms_per_step, steps, n, ...)."""
    import numpy as np
    note = _pick_ref_subconfig()
    ref = _ref_lib()
    cores = _host_cores()
    ref.set_num_threads(cores)
    rng = np.random.default_rng(0xB200)

    def run(n):
        if op == "dgemm":
            a = np.asfortranarray(rng.uniform(-1, 1, (n, n))); b = np.asfortranarray(rng.uniform(-1, 1, (n, n)))
            c = np.asfortranarray(rng.uniform(-1, 1, (n, n)))
            t0 = time.perf_counter(); ref.gemm(0, 0, ALPHA, a, b, BETA, c); dt = time.perf_counter() - t0
            return 2.0 * n ** 3, dt
        m, nn = 2 * n, max(1, n // 2)                            # T1 aspect: m = 32768, n = 8192 at n = 16384
        a = np.tril(rng.uniform(-1, 1, (m, m)) / np.sqrt(m)); a[np.arange(m), np.arange(m)] += 2.0
        a = np.asfortranarray(a)
        b = np.asfortranarray(rng.uniform(-1, 1, (m, nn)))
        t0 = time.perf_counter(); ref.trsm(0, 0xC0, 0, 0, ALPHA, a, b); dt = time.perf_counter() - t0
        return 1.0 * m * m * nn, dt

    fl, dt = run(1024 if op == "dtrsm" else 2048)                # warm-up (thread pool, pack buffers) + speed estimate
    fl, dt = run(1024 if op == "dtrsm" else 2048)
    rate = fl / dt
    n = n_full
    flops_full = 2.0 * float(n_full) ** 3 if op == "dgemm" else 1.0 * (2.0 * n_full) ** 2 * (n_full // 2)
    # small sizes under-estimate the large-size rate, so the prediction is conservative
    steps = max_steps
    while steps > 1 and flops_full / rate * steps > budget_s:
        steps -= 1
    f = flops_full
    while n > 2048 and f / rate * steps > budget_s:
        n //= 2
        f /= 8.0
    tot_f = tot_t = 0.0
    best = 0.0
    for _ in range(steps):
        fl, dt = run(n)
        tot_f += fl; tot_t += dt; best = max(best, fl / dt)
    shape = f"m=n=k={n}" if op == "dgemm" else f"m={2 * n} n={max(1, n // 2)}"
    return {"value": tot_f / tot_t / 1e9, "best": best / 1e9, "ms_per_step": tot_t / steps * 1e3, "steps": steps, "n": n, "cores": cores,
            "full_size": n == n_full,
            "sample": f"{op} {shape} column-major, {steps} timed step(s), reference BLIS sub-config '{ref.arch()}' [{note}], {cores} threads"}


def reference_arm(args, workload: str) -> int:
    if args.arch_probe:
        return _arch_probe()
    budget = float(os.environ.get("B200_REF_BUDGET", "150"))
    out = {}
    ops = ("dgemm", "dtrsm") if args.op == "both" else (args.op,)
    for op in ops:
        out[op] = cpu_reference(op, budget_s=budget if op == "dgemm" else budget / 2, max_steps=max(1, min(args.steps, 3)) if op == "dgemm" else 1,
                                n_full=args.n)
    if args.cpu_sample:                                          # internal: the b200 arm's cpu_baseline leg
        print(json.dumps(out))
        return 0
    head = out.get("dgemm") or out["dtrsm"]
    line = {
        "impl": "reference", "metric": METRIC, "value": head["value"], "unit": "GFLOPS", "n_gpus": args.gpus,
        "steps": head["steps"], "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "full_size": head["full_size"],
                   "note": "the reference's own CPU implementation on the host cores; a step is one call at the size named in cpu_baseline.sample "
                           "(the full workload unless the predicted time exceeded the budget)"},
        "cpu_baseline": {"value": head["value"], "unit": "GFLOPS", "cores": head["cores"], "kind": "reference", "sample": head["sample"]},
        "e2e": {"value": head["value"], "unit": "GFLOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if "dtrsm" in out and "dgemm" in out:
        t = out["dtrsm"]
        line["dtrsm"] = {"value": t["value"], "unit": "GFLOPS", "ms_per_step": t["ms_per_step"], "full_size": t["full_size"],
                         "cpu_baseline": {"value": t["value"], "unit": "GFLOPS", "cores": t["cores"], "kind": "reference", "sample": t["sample"]}}
    print(json.dumps(line))
    return 0


def cpu_baseline_subprocess(op: str, budget: float) -> dict | None:
    """The reference on the host cores in its OWN process (this one may have the B200 plugin registered)."""
    env = dict(os.environ, B200_REF_BUDGET=str(budget))
    env.pop("RANK", None); env.pop("WORLD_SIZE", None); env.pop("LOCAL_RANK", None)
    try:
        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--cpu-sample", "--op", op, "--steps", "1"],
                           env=env, capture_output=True, text=True, timeout=600)
        if r.returncode != 0 or not r.stdout.strip():
            print(f"cpu baseline failed: {r.stderr[-500:]}", file=sys.stderr)
            return None
        return json.loads(r.stdout.strip().splitlines()[-1])
    except (subprocess.TimeoutExpired, ValueError, OSError) as exc:
        print(f"cpu baseline failed: {exc!r}", file=sys.stderr)
        return None


# ----------------------------------------------------------------------------- synthetic inputs
def rnd_colmajor(torch, rows, cols, device, gen):
    """rows x cols column-major, uniform[-1,1] divided by the 1-norm rounded up to a power of two
    (libblis_test_mobj_randomize, testsuite/src/test_libblis.c:2529-2565)."""
    x = torch.rand(cols, rows, dtype=torch.float64, device=device, generator=gen) * 2 - 1
    nrm = float(x.abs().sum(dim=1).max())
    return (x / float(2 ** math.ceil(math.log2(nrm)))).t()


def trsm_inputs(torch, m, n_cols, device, seed, b_seed=None):
    """T1 inputs: A random, normalised, diag += 2.0 (test_libblis.c:2583-2589), column-major; B column-major m x n_cols."""
    g = torch.Generator(device=device); g.manual_seed(seed)
    at = torch.rand(m, m, dtype=torch.float64, device=device, generator=g) * 2 - 1
    at = at / float(at.abs().sum(dim=1).max()); at.diagonal().add_(2.0)
    at = at.t()                                                   # column-major; the lower triangle is what trsm reads
    if b_seed is not None:
        g.manual_seed(b_seed)
    bt0 = (torch.rand(n_cols, m, dtype=torch.float64, device=device, generator=g) * 2 - 1).t()
    return at, bt0


def measured_hbm_peak():
    """(GB/s, source): the driver-written copy bandwidth of this pool's B200s, else the profiling recipe's fallback."""
    try:
        v = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
        return v, "MEASURED_PEAKS.json hbm_gbs (driver-written: torch copy of 1 Gi bf16 elements, read+write bytes)"
    except (OSError, ValueError, KeyError):
        return 6650.0, "fallback of /opt/skills/guides/B200_PROFILING.md (no MEASURED_PEAKS.json)"


def gemm_kernel_of(stats: dict, by_flops: bool = False) -> str:
    """The gemm tile kernel with the most launches in a kernel histogram.  by_flops (dtrsm): the histogram of a solve is
    dominated in COUNT by the small updates (k <= 256, the D-staging variant CST=1, under 1 % of the flops: level k of the
    recursion holds k/m of them) and the split-k tails, so the kernel that bounds the step is looked for among the plain
    variants first."""
    g = {k: v for k, v in stats.items() if k.startswith("gemm_")}
    if by_flops:
        big = {k: v for k, v in g.items() if "CST=1" not in k and "SK=1" not in k}
        g = big or g
    return max(g, key=g.get) if g else (max(stats, key=stats.get) if stats else "none")


# ----------------------------------------------------------------------------- e2e through the reference's BLAS layer
class RefFrontEnd:
    """The reference's Fortran BLAS entry points (dgemm_, dtrsm_: frame/compat/bla_gemm.c:127-259, bla_trsm.c:126-217) served
    by the engine: blis_glue's plugin is registered on the unmodified reference library (INTEGRATION.md route 1)."""

    _inst = None

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self):
        if not (GLUE_SO.exists() and REF_SO.exists()):
            raise FileNotFoundError("oracle/_ref/libblis_b200_glue.so or libblis_ref.so not built")
        self.glue = ctypes.CDLL(str(GLUE_SO), mode=ctypes.RTLD_GLOBAL)     # first: its bli_*_ex definitions interpose
        self.ref = ctypes.CDLL(str(REF_SO), mode=ctypes.RTLD_GLOBAL)
        self.glue.bli_plugin_register_b200.restype = ctypes.c_int
        if self.glue.bli_plugin_register_b200() != -1:
            raise RuntimeError("bli_plugin_register_b200 failed")

    @staticmethod
    def _i(v): return ctypes.byref(ctypes.c_int(int(v)))
    @staticmethod
    def _d(v): return ctypes.byref(ctypes.c_double(float(v)))

    def dgemm_(self, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc):
        self.ref.dgemm_(ctypes.c_char_p(b"N"), ctypes.c_char_p(b"N"), self._i(m), self._i(n), self._i(k), self._d(alpha),
                        ctypes.c_void_p(a), self._i(lda), ctypes.c_void_p(b), self._i(ldb), self._d(beta), ctypes.c_void_p(c), self._i(ldc))

    def dtrsm_(self, m, n, alpha, a, lda, b, ldb):
        self.ref.dtrsm_(ctypes.c_char_p(b"L"), ctypes.c_char_p(b"L"), ctypes.c_char_p(b"N"), ctypes.c_char_p(b"N"), self._i(m), self._i(n),
                        self._d(alpha), ctypes.c_void_p(a), self._i(lda), ctypes.c_void_p(b), self._i(ldb))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--op", default="both", choices=["both", "dgemm", "dtrsm"])
    ap.add_argument("--n", type=int, default=N_DEFAULT)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-peak", action="store_true", help="skip the FP64 peak microbenchmark (profiling runs)")
    ap.add_argument("--no-check", action="store_true", help="N > 1: skip the post-timing parity check of the distributed result")
    ap.add_argument("--no-g3", action="store_true", help="N = 8: skip the 65536^3 leg (BASELINE configs[4])")
    ap.add_argument("--no-skinny", action="store_true", help="skip the skinny k=64 leg (BASELINE configs[2] shape, HBM roofline)")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling leg (one 16384^3 over all GPUs)")
    ap.add_argument("--workload", default="headline", choices=["headline", "g3"],
                    help="g3 = BASELINE configs[4]: dgemm m=n=k=65536 2D-sharded over all ranks")
    ap.add_argument("--arch-probe", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--cpu-sample", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.n
    wl_gemm = f"dgemm m=n=k={n} column-major device-resident fp64, alpha={ALPHA} beta={BETA}"
    wl_trsm = f"dtrsm left/lower/notrans/nonunit m={2 * n} n={n // 2} column-major fp64, alpha={ALPHA}"
    workload = wl_trsm if args.op == "dtrsm" else wl_gemm

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, workload)

    # ------------------------------------------------------------------ b200 arm
    # stdout carries ONE JSON line (driver contract).  Libraries write to fd 1 behind Python's back -- NCCL prints its
    # version banner there when the engine creates its communicator (N > 1) --, so fd 1 points at stderr for the whole run
    # and the line leaves through a saved duplicate of the real stdout.
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj: dict) -> None:
        os.write(out_fd, (json.dumps(obj) + "\n").encode())

    import torch
    from blis_b200 import api
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # the panel all-gathers move ~10 GB/s per GPU: a few NCCL channels are plenty; every channel is an SM
        # taken from the DMMA kernel for a moment, which its dynamic tile scheduler absorbs (no SMs reserved)
        nch = int(os.environ.get("B200_NCCL_CHANNELS", "8"))
        os.environ.setdefault("NCCL_MAX_NCHANNELS", str(nch))
        os.environ.setdefault("NCCL_MIN_NCHANNELS", str(nch))
        dist.init_process_group("nccl", device_id=dev)
        api.set_option("reserve_sms", int(os.environ.get("B200_RESERVE_SMS", "0")))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.Stream(device=dev)
    peak = None

    def dmma_peak():
        nonlocal peak
        if peak is None:
            peak = 37.08 if args.no_peak else api.measure_peak("dmma", 300)     # TFLOP/s, FP64 tensor pipe, this GPU, this run
        return peak

    PEAK_SRC = ("FP64 DMMA (mma.sync m8n8k4) microbenchmark measured in this run by b200_measure_peak; "
                "MEASURED_PEAKS.json holds only HBM GB/s and bf16 TFLOP/s")

    def timed(step, steps, warmup, around=None):
        """W warm-ups, then K steps bracketed by barrier + synchronize; CUDA events on the launching stream.
        Returns (ms per step as max over ranks of the whole region / K, per-step ms list, launches, kernel histogram, clocks)."""
        for _ in range(warmup):
            step()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        api.kernel_stats(reset=True)
        l0 = api.launch_count()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for i in range(steps):
            if around is not None:
                around()                                         # untimed per-step preparation (restore B for trsm)
            ev[i][0].record(stream)
            step()
            ev[i][1].record(stream)
        barrier()
        launches = api.launch_count() - l0
        stats = api.kernel_stats()
        clocks = sampler.stop() if rank == 0 else None
        per = [a.elapsed_time(b) for a, b in ev]
        if around is None:
            total = ev[0][0].elapsed_time(ev[-1][1])             # contiguous region: EXACTLY K steps
        else:
            total = sum(per)
        return allmax(total) / steps, per, launches, stats, clocks

    result = {}
    with torch.cuda.stream(stream):
        # ============================================================== dgemm
        gemm_line = None
        if args.op in ("both", "dgemm"):
            if world == 1:
                g = torch.Generator(device=dev); g.manual_seed(0xB200 + rank)
                a, b, c = (rnd_colmajor(torch, n, n, dev, g) for _ in range(3))

                def step():
                    api.bli_dgemm(0, 0, n, n, n, ALPHA, a, 1, n, b, 1, n, BETA, c, 1, n)
                total_flops = flops_rank = 2.0 * n ** 3
                parallelism, job = "single GPU", None
            else:
                from blis_b200 import dist as bdist
                kb = int(os.environ.get("B200_DIST_KB", "2048"))
                if args.workload == "g3":
                    job = bdist.DistGemm(65536, 65536, 65536, world, rank, dev, alpha=ALPHA, beta=BETA, kb=kb)
                    wl_gemm = workload = "dgemm m=n=k=65536 column-major fp64, 2D-sharded, alpha=2.0 beta=1.2 (BASELINE configs[4])"
                else:
                    job = bdist.WeakScalingGemm(n, world, rank, dev, alpha=ALPHA, beta=BETA, kb=kb)
                step = job.step
                total_flops, flops_rank = job.total_flops, job.total_flops / world
                parallelism = job.describe()
            ms, per, launches, stats, clocks = timed(step, args.steps, args.warmup)
            value = total_flops / (ms * 1e-3) / 1e9
            wait_ms = transport = None
            if world > 1 and job.native:
                # how long the compute stream sat waiting for gathers in one more product (events around every wait, in the engine);
                # taken right behind the timed region, while all ranks are still in step
                job.step(flags=api.DIST_AB_STATIC | api.DIST_TRACE)
                wait_ms = allmax(api.dist_last_wait_ms())
                transport = api.dist_transport()
            roof = None
            if rank == 0:
                kern_ms = sum(per) / len(per)
                achieved = flops_rank / (kern_ms * 1e-3) / 1e12
                kern = gemm_kernel_of(stats)
                traffic = None
                tfile = ROOT / "profiles" / "dgemm_traffic.json"
                if world == 1 and n == N_DEFAULT and kern.startswith("gemm_dmma_tma_kernel") and tfile.exists():
                    try:
                        traffic = json.loads(tfile.read_text()).get("dram_bytes_per_launch")   # ncu --set full capture of this kernel at this size
                    except (ValueError, OSError):
                        traffic = None
                roof = {"bound": "tensor", "achieved": achieved, "peak": dmma_peak(), "unit": "TFLOP/s", "frac": achieved / dmma_peak(),
                        "traffic": traffic, "peak_source": PEAK_SRC, "kernel": kern, "kernels": stats,
                        "algorithmic_flop_per_step_per_gpu": flops_rank,
                        "algorithmic_bytes_per_step_per_gpu": 8.0 * 4 * n * n if world == 1 else None}
            if roof is not None and wait_ms is not None:
                roof["gather_wait_ms_per_product_max_over_ranks"] = wait_ms
                roof["panel_transport"] = transport
            gemm_line = dict(value=value, ms=ms, launches=launches, roof=roof, clocks=clocks, parallelism=parallelism, total_flops=total_flops)

            # ---- e2e
            e2e = None
            if not args.no_e2e and world == 1:
                ah, bh, ch = (torch.empty(n, n, dtype=torch.float64).pin_memory().t() for _ in range(3))
                ah.copy_(a); bh.copy_(b); ch.copy_(c)
                torch.cuda.synchronize()
                try:
                    fe = RefFrontEnd.get()
                    api._lib.load().b200_set_stream(None)        # the BLAS layer runs on the engine's own stream
                    call = lambda: fe.dgemm_(n, n, n, ALPHA, ah.data_ptr(), n, bh.data_ptr(), n, BETA, ch.data_ptr(), n)   # noqa: E731
                    how = ("the reference's Fortran entry point dgemm_ (frame/compat/bla_gemm.c:127-259, unmodified libblis) with the B200 plugin "
                           "registered -> bli_gemm_ex -> blis_glue -> b200_gemm, on pinned host operands")
                except (OSError, RuntimeError, FileNotFoundError) as exc:
                    call = lambda: api.bli_dgemm(0, 0, n, n, n, ALPHA, ah, 1, n, bh, 1, n, BETA, ch, 1, n)                # noqa: E731
                    how = f"b200_gemm through the ctypes mirror (reference front end unavailable: {exc}) on pinned host operands"
                e2e_steps = max(2, min(args.steps, 3))
                l0 = api.launch_count()
                call()                                            # warm-up (allocations)
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    call()                                        # returns after C is back in host memory
                dt = (time.perf_counter() - t0) / e2e_steps
                e2e = {"value": 2.0 * n ** 3 / dt / 1e9, "unit": "GFLOPS", "h2d_bytes_per_step": 3 * n * n * 8,
                       "d2h_bytes_per_step": n * n * 8, "ms_per_step": dt * 1e3, "engine_launches": api.launch_count() - l0,
                       "how": how + ": H2D of A,B,C + kernels + D2H of C per step, host wall clock"}
                del ah, bh, ch
            hosts = None
            if not args.no_e2e and world > 1 and args.workload == "headline":
                e2e, hosts = dist_e2e(args, job, dist, dev, rank, barrier, torch)
            gemm_line["e2e"] = e2e

            # ---- N > 1: parity check of the distributed result (device shards, and host shards when e2e ran)
            check = None
            if world > 1 and not args.no_check:
                ck = job.verify(hosts=hosts)
                t = torch.tensor([1.0 if ck["bit_equal"] else 0.0, 1.0 if ck.get("host_bit_equal", True) else 0.0, -ck["resid"]],
                                 dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                check = {"bit_equal_to_single_gpu_replay": bool(t[0].item() == 1.0), "testsuite_resid_max_over_ranks": float(-t[2].item()),
                         "resid_pass_threshold": 1e-14, "how": ck["how"]}
                if "host_bit_equal" in ck:
                    check["host_shards_bit_equal_to_device_shards"] = bool(t[1].item() == 1.0)
            gemm_line["check"] = check
            del hosts

            # ---- strong scaling: ONE n^3 product over all GPUs
            strong = None
            if world > 1 and not args.no_strong and args.workload == "headline":
                from blis_b200 import dist as bdist
                job.close()
                del job
                torch.cuda.empty_cache()
                sjob = bdist.DistGemm(n, n, n, world, rank, dev, alpha=ALPHA, beta=BETA, kb=int(os.environ.get("B200_DIST_KB_STRONG", "1024")))
                sms, sper, _, sstats, _ = timed(sjob.step, args.steps, args.warmup)
                scheck = None if args.no_check else sjob.verify()
                if scheck is not None:
                    t = torch.tensor([1.0 if scheck["bit_equal"] else 0.0, -scheck["resid"]], dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MIN)
                    scheck = {"bit_equal_to_single_gpu_replay": bool(t[0].item() == 1.0), "testsuite_resid_max_over_ranks": float(-t[1].item())}
                strong = {"value": sjob.total_flops / (sms * 1e-3) / 1e9, "unit": "GFLOPS", "ms_per_step": sms, "scaling": "strong",
                          "workload": wl_gemm, "parallelism": sjob.describe(), "kernel": gemm_kernel_of(sstats), "check": scheck}
                sjob.close()
                del sjob
            gemm_line["strong"] = strong

            # ---- BASELINE configs[4]: dgemm 65536^3 2D-sharded over 8 GPUs (only at N = 8; strong-scaled by definition)
            g3 = None
            if world == 8 and not args.no_g3 and args.workload == "headline":
                from blis_b200 import dist as bdist
                job = None
                torch.cuda.empty_cache()
                gjob = bdist.DistGemm(65536, 65536, 65536, world, rank, dev, alpha=ALPHA, beta=BETA, kb=kb)
                gms, gper, _, gstats, _ = timed(gjob.step, 2, 3)
                gcheck = None if args.no_check else gjob.verify()
                if gcheck is not None:
                    t = torch.tensor([1.0 if gcheck["bit_equal"] else 0.0, -gcheck["resid"]], dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MIN)
                    gcheck = {"bit_equal_to_single_gpu_replay": bool(t[0].item() == 1.0), "testsuite_resid_max_over_ranks": float(-t[1].item())}
                g3 = {"value": gjob.total_flops / (gms * 1e-3) / 1e9, "unit": "GFLOPS", "ms_per_step": gms, "steps": 2, "warmup": 3, "scaling": "strong",
                      "workload": "dgemm m=n=k=65536 column-major fp64, 2D-sharded over 8 GPUs, alpha=2.0 beta=1.2 (BASELINE configs[4])",
                      "parallelism": gjob.describe(), "kernel": gemm_kernel_of(gstats), "check": gcheck,
                      "frac_of_dmma_peak_per_gpu": gjob.total_flops / world / (gms * 1e-3) / 1e12 / dmma_peak()}
                gjob.close()
                del gjob
                torch.cuda.empty_cache()
            gemm_line["g3"] = g3
            if world == 1:
                del a, b, c
            torch.cuda.empty_cache()

        # ============================================================== skinny dgemm (BASELINE configs[2]: the sup shapes, k = 64)
        skinny_line = None
        if args.op == "both" and args.workload == "headline" and not args.no_skinny:
            from blis_b200 import dist as bdist
            ks = 64
            sj = bdist.DistSkinnyGemm(n, n * world, ks, world, rank, dev, alpha=ALPHA, beta=BETA, root=0)
            ssteps = max(args.steps, 20)
            sms, sper, sl, sstats, _ = timed(sj.step, ssteps, args.warmup)
            sck = None if args.no_check else sj.verify()
            if sck is not None:
                t = torch.tensor([1.0 if sck["bit_equal"] else 0.0, -sck["resid"]], dtype=torch.float64, device=dev)
                if dist is not None:
                    dist.all_reduce(t, op=dist.ReduceOp.MIN)
                sck = {"bit_equal_to_single_gpu_engine": bool(t[0].item() == 1.0), "testsuite_resid_max_over_ranks": float(-t[1].item())}
            if rank == 0:
                hbm = measured_hbm_peak()
                gbs = sj.bytes_rank / (sum(sper) / len(sper) * 1e-3) / 1e9
                skinny_line = {"value": sj.total_flops / (sms * 1e-3) / 1e9, "unit": "GFLOPS", "ms_per_step": sms, "steps": ssteps,
                               "scaling": "weak", "workload": f"dgemm m={n} n={n}*N k={ks} column-major fp64, alpha=2.0 beta=1.2 (BASELINE configs[2] skinny shape per GPU)",
                               "parallelism": sj.describe() if world > 1 else "single GPU", "gpu_launches": sl, "check": sck,
                               "l2": "C block (2 GiB per GPU) far exceeds the 126 MB L2; A and B (8 MB each) are L2 resident by design",
                               "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm[0], "unit": "GB/s", "frac": gbs / hbm[0], "traffic": None,
                                            "peak_source": hbm[1], "kernel": gemm_kernel_of(sstats), "kernels": sstats,
                                            "algorithmic_bytes_per_step_per_gpu": sj.bytes_rank,
                                            "dmma_frac": (sj.total_flops / world) / (sum(sper) / len(sper) * 1e-3) / 1e12 / dmma_peak()}}
            del sj
            torch.cuda.empty_cache()

        # ============================================================== dtrsm
        trsm_line = None
        if args.op in ("both", "dtrsm") and args.workload == "headline":
            m_t, n_t = 2 * n, n // 2
            if world == 1:
                j0, j1 = 0, n_t
                par_t = "single GPU"
            else:
                from blis_b200 import dist as bdist
                j0, j1 = bdist.trsm_column_block(rank, world, n_t)
                par_t = (f"B split into {world} column blocks (bli_thread_range_sub, bf=128), A replicated; strong scaling, no data-path collective; "
                     "one b200_dist_trsm call per solve")
            at, bt0 = trsm_inputs(torch, m_t, j1 - j0, dev, 0xB200, b_seed=0xB201 + rank)
            bt = bt0.clone(memory_format=torch.preserve_format)

            def restore():
                bt.copy_(bt0)

            if world > 1:
                bdist.native_init(dev)

            def tstep():
                if world > 1:       # ONE C-ABI call: the engine cuts this rank's column block itself (b200_dist_trsm; A already replicated)
                    api.dist_trsm(torch.float64, 0, 0xC0, 0, 0, -1, m_t, n_t, ALPHA, at, 1, m_t, bt, 1, m_t)
                else:
                    api.bli_dtrsm(0, 0xC0, 0, 0, m_t, j1 - j0, ALPHA, at, 1, m_t, bt, 1, m_t)
            restore()
            tsteps = max(2, min(args.steps, 5))
            tms, tper, tl, tstats, tclocks = timed(tstep, tsteps, args.warmup, around=restore)
            tflops = 1.0 * m_t * m_t * n_t
            tval = tflops / (tms * 1e-3) / 1e9
            troof = None
            if rank == 0:
                ach = (1.0 * m_t * m_t * (j1 - j0)) / (sum(tper) / len(tper) * 1e-3) / 1e12
                troof = {"bound": "tensor", "achieved": ach, "peak": dmma_peak(), "unit": "TFLOP/s", "frac": ach / dmma_peak(), "traffic": None,
                         "peak_source": PEAK_SRC, "kernel": gemm_kernel_of(tstats, by_flops=True), "kernels": tstats,
                         "algorithmic_flop_per_step_per_gpu": 1.0 * m_t * m_t * (j1 - j0),
                         "algorithmic_bytes_per_step_per_gpu": 8.0 * (m_t * m_t / 2 + 2 * m_t * (j1 - j0))}
            # parity of the timed configuration: the testsuite residual of the last solve (test_trsm.c:362-381 form)
            g2 = torch.Generator(device=dev); g2.manual_seed(11)
            tvec = (torch.rand(j1 - j0, dtype=torch.float64, device=dev, generator=g2) * 2 - 1) / n_t
            xt, rhs = bt @ tvec, ALPHA * (bt0 @ tvec)
            w = torch.linalg.solve_triangular(at, rhs.unsqueeze(1), upper=False).squeeze(1)
            tresid = allmax(float(torch.linalg.vector_norm(xt - w) / max(1.0, float(torch.linalg.vector_norm(w)))))
            te2e = None
            if not args.no_e2e and world == 1:
                try:
                    fe = RefFrontEnd.get()
                    ah = torch.empty(m_t, m_t, dtype=torch.float64).pin_memory().t(); ah.copy_(at)
                    bh0 = torch.empty(n_t, m_t, dtype=torch.float64).pin_memory().t(); bh0.copy_(bt0)
                    bh = torch.empty(n_t, m_t, dtype=torch.float64).pin_memory().t()
                    torch.cuda.synchronize()
                    api._lib.load().b200_set_stream(None)
                    bh.copy_(bh0); fe.dtrsm_(m_t, n_t, ALPHA, ah.data_ptr(), m_t, bh.data_ptr(), m_t)           # warm-up
                    tt = 0.0
                    for _ in range(2):
                        bh.copy_(bh0)
                        t0 = time.perf_counter(); fe.dtrsm_(m_t, n_t, ALPHA, ah.data_ptr(), m_t, bh.data_ptr(), m_t); tt += time.perf_counter() - t0
                    dt = tt / 2
                    te2e = {"value": tflops / dt / 1e9, "unit": "GFLOPS", "ms_per_step": dt * 1e3,
                            "h2d_bytes_per_step": int(8 * (m_t * (m_t + 1024) // 2 + m_t * n_t)), "d2h_bytes_per_step": 8 * m_t * n_t,
                            "how": "the reference's dtrsm_ (frame/compat/bla_trsm.c:126-217) with the B200 plugin registered, pinned host A and B: H2D of "
                                   "B in row blocks and of the stored triangle of A (per block row the block its update gemm reads, then diagonal squares "
                                   "of <= 1024 rows, in the order the solve reads them, under the solve), kernels, D2H of X block row by block row "
                                   "as they are solved (host_trsm.cuh: trsm_host_rowpipe); host wall clock"}
                    del ah, bh, bh0
                except (OSError, RuntimeError, FileNotFoundError) as exc:
                    print(f"dtrsm e2e skipped: {exc!r}", file=sys.stderr)
            trsm_line = {"value": tval, "unit": "GFLOPS", "ms_per_step": tms, "steps": tsteps, "scaling": "strong" if world > 1 else "weak",
                         "workload": wl_trsm, "parallelism": par_t, "gpu_launches": tl, "roofline": troof, "clocks": tclocks,
                         "testsuite_resid": tresid, "e2e": te2e}
            del at, bt, bt0
            torch.cuda.empty_cache()

    # ------------------------------------------------------------------ CPU baseline (separate process: no plugin there)
    cpu = cpu_t = None
    if rank == 0 and not args.no_cpu and world == 1:
        both = cpu_baseline_subprocess("both" if args.op == "both" else args.op, budget=30.0)
        if both:
            if "dgemm" in both:
                d = both["dgemm"]; cpu = {"value": d["value"], "unit": "GFLOPS", "cores": d["cores"], "kind": "reference", "sample": d["sample"]}
            if "dtrsm" in both:
                d = both["dtrsm"]; cpu_t = {"value": d["value"], "unit": "GFLOPS", "cores": d["cores"], "kind": "reference", "sample": d["sample"]}

    if rank == 0:
        head = gemm_line if gemm_line is not None else None
        if trsm_line is not None:
            trsm_line["cpu_baseline"] = cpu_t
        if head is not None:
            line = {
                "metric": METRIC, "value": head["value"], "unit": "GFLOPS", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": head["ms"], "higher_is_better": True,
                "scaling": "strong" if (args.workload == "g3" and world > 1) else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "parallelism": head["parallelism"],
                           "l2": "inputs (3 x 2 GiB per GPU) far exceed the 126 MB L2; no flush needed",
                           "timing": "CUDA events on the launching stream, max over ranks"},
                "clocks": head["clocks"], "e2e": head["e2e"], "gpu_launches": head["launches"], "roofline": head["roof"], "cpu_baseline": cpu,
                "dtrsm": trsm_line, "skinny": skinny_line,
            }
            if world > 1:
                line["check"] = head["check"]; line["strong"] = head["strong"]; line["g3"] = head.get("g3")
        else:
            t = trsm_line
            line = {
                "metric": METRIC, "value": t["value"], "unit": "GFLOPS", "n_gpus": world, "steps": t["steps"], "warmup": args.warmup,
                "ms_per_step": t["ms_per_step"], "higher_is_better": True, "scaling": t["scaling"], "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": wl_trsm, "parallelism": t["parallelism"],
                                                "l2": "A (8.6 GB) and B (2 GiB) far exceed the 126 MB L2; no flush needed",
                                                "timing": "CUDA events on the launching stream around every solve, max over ranks"},
                "clocks": t["clocks"], "e2e": t["e2e"], "gpu_launches": t["gpu_launches"], "roofline": t["roofline"], "cpu_baseline": cpu_t,
                "testsuite_resid": t["testsuite_resid"],
            }
        emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def dist_e2e(args, job, dist, dev, rank, barrier, torch):
    """e2e at N > 1: every rank keeps its shards (block-cyclic k panels of A and B, its block of C) in pinned HOST memory;
    a step is DistGemm.step_host: shards uploaded in the order the k steps use them, C in column blocks under the first
    k step, finished column blocks of C read back under the last one (blis_b200/dist.py: summa_host).  All ranks agree
    first that their pinned buffers exist, so that no rank can leave the others inside a collective."""
    try:
        hosts, ok = job.host_shards(), 1
    except Exception as exc:                                      # noqa: BLE001 (pinning can fail on a small host)
        print(f"[rank {rank}] e2e skipped: {exc}", file=sys.stderr)
        hosts, ok = None, 0
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) != 1:
        return None, None
    try:
        h2d = sum(h.numel() * h.element_size() for h in hosts)
        d2h = hosts[2].numel() * hosts[2].element_size()

        def e2e_step():
            job.step_host(hosts)
            torch.cuda.synchronize()                              # C is back in host memory
        e2e_steps = max(2, min(args.steps, 3))
        e2e_step()                                                # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        tt = torch.tensor([(time.perf_counter() - t0) / e2e_steps, float(h2d), float(d2h)], dtype=torch.float64, device=dev)
        tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        dt = float(tmax[0].item())
        return {"value": job.total_flops / dt / 1e9, "unit": "GFLOPS", "h2d_bytes_per_step": int(tt[1].item()),
                "d2h_bytes_per_step": int(tt[2].item()), "ms_per_step": dt * 1e3,
                "how": "DistGemm.step_host: every rank's shards of A, B and C live in pinned host memory; H2D in k-step order, "
                       "C in column blocks under the first k step, C blocks read back under the last; host wall clock "
                       "between barriers, max over ranks; bytes summed over ranks"}, hosts
    except Exception as exc:                                      # noqa: BLE001 -- keep the device-resident line
        print(f"[rank {rank}] e2e failed: {exc!r}", file=sys.stderr)
        return None, None


if __name__ == "__main__":
    sys.exit(main())
