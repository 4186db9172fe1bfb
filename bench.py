#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 level-3 engine (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--op dgemm|dtrsm]

Workload (BASELINE.json configs[1]): dgemm m=n=k=16384, column-major,
alpha=2.0, beta=1.2 (testsuite/src/test_gemm.c:213-214), synthetic inputs
uniform[-1,1] normalised as libblis_test_mobj_randomize does.  One "step" is one
dgemm over that batch.  N>1 (torchrun): weak scaling -- every rank owns a
16384x16384 block of C of a larger product, partitioned as a 2D block
decomposition with NCCL all-gather of the A/B k-panels (blis_b200/dist.py).

JSON line keys follow the contract: `value` is device-resident throughput,
`e2e` is the same call with HOST (pinned) operands -- H2D of A,B,C and D2H of C
inside the timed region --, `roofline` is the dgemm kernel against the FP64
tensor-pipe peak measured in the same run, `cpu_baseline` is the real reference
BLIS (oracle/_ref, test infrastructure) timed on the host cores on a bounded
sample.  `--impl reference` times only that CPU reference.
"""
from __future__ import annotations

import argparse
import json
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
def _ref_lib():
    sys.path.insert(0, str(ROOT / "tests"))
    from refblis import RefBlis, have_ref
    if not have_ref():
        sys.path.insert(0, str(ROOT / "oracle"))
        import build_ref
        build_ref.build()
    return RefBlis()


def cpu_reference_gflops(op: str, budget_s: float = 20.0, reps: int = 2):
    """Reference BLIS (oracle/_ref, all host threads) on a bounded sample of the workload."""
    import numpy as np
    ref = _ref_lib()
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        pass
    ref.set_num_threads(cores)
    rng = np.random.default_rng(0xB200)

    def run(n):
        if op == "dgemm":
            a = np.asfortranarray(rng.uniform(-1, 1, (n, n))); b = np.asfortranarray(rng.uniform(-1, 1, (n, n)))
            c = np.asfortranarray(rng.uniform(-1, 1, (n, n)))
            t0 = time.perf_counter(); ref.gemm(0, 0, ALPHA, a, b, BETA, c); dt = time.perf_counter() - t0
            return 2.0 * n ** 3 / dt / 1e9
        m, nn = n, max(1, n // 4)                              # same 4:1 aspect as m=32768, n=8192
        a = np.asfortranarray(np.tril(rng.uniform(-1, 1, (m, m)) / np.sqrt(m)) + 2.0 * np.eye(m))
        b = np.asfortranarray(rng.uniform(-1, 1, (m, nn)))
        t0 = time.perf_counter(); ref.trsm(0, 0xC0, 0, 0, ALPHA, a, b); dt = time.perf_counter() - t0
        return 1.0 * m * m * nn / dt / 1e9

    n = 2048
    g = run(n)                                                    # warm-up + speed estimate
    flops_per_n3 = 2.0 if op == "dgemm" else 0.25
    while n < N_DEFAULT and flops_per_n3 * (2 * n) ** 3 / (g * 1e9) * reps < budget_s:
        n *= 2
    best = max(run(n) for _ in range(reps))
    shape = f"m=n=k={n}" if op == "dgemm" else f"m={n} n={max(1, n // 4)}"
    return best, cores, f"{op} {shape} column-major, best of {reps}, reference BLIS sub-config '{ref.arch()}', {cores} threads", n


# ----------------------------------------------------------------------------- synthetic inputs
def make_inputs(torch, n, device, seed):
    g = torch.Generator(device=device); g.manual_seed(seed)

    def rnd():
        x = torch.rand(n, n, dtype=torch.float64, device=device, generator=g) * 2 - 1
        nrm = float(x.abs().sum(dim=1).max())
        import math
        return (x / float(2 ** math.ceil(math.log2(nrm)))).t()      # column-major view: strides (1, n)
    return rnd(), rnd(), rnd()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--op", default="dgemm", choices=["dgemm", "dtrsm"])
    ap.add_argument("--n", type=int, default=N_DEFAULT)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-peak", action="store_true", help="skip the FP64 peak microbenchmark (profiling runs)")
    ap.add_argument("--workload", default="headline", choices=["headline", "g3"],
                    help="g3 = BASELINE configs[4]: dgemm m=n=k=65536 2D-sharded over all ranks")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n = args.n
    workload = (f"dgemm m=n=k={n} column-major device-resident fp64, alpha={ALPHA} beta={BETA}"
                if args.op == "dgemm" else f"dtrsm left/lower/notrans/nonunit m={2 * n} n={n // 2} column-major fp64, alpha={ALPHA}")

    # ------------------------------------------------------------------ reference arm: CPU only, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return 0
        vals = []
        info = None
        for _ in range(max(1, min(args.steps, 3))):
            v, cores, sample, _ = cpu_reference_gflops(args.op, budget_s=12.0, reps=1)
            vals.append(v); info = (cores, sample)
        v = max(vals)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": "GFLOPS", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "note": "each step is a bounded sample of the workload on the host CPU"},
            "cpu_baseline": {"value": v, "unit": "GFLOPS", "cores": info[0], "kind": "reference", "sample": info[1]},
            "e2e": {"value": v, "unit": "GFLOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    # ------------------------------------------------------------------ b200 arm
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

    stream = torch.cuda.Stream(device=dev)
    launches = [0]
    with torch.cuda.stream(stream):
        if world == 1:
            a, b, c = make_inputs(torch, n, dev, 0xB200 + rank)
            if args.op == "dgemm":
                def step():
                    api.bli_dgemm(0, 0, n, n, n, ALPHA, a, 1, n, b, 1, n, BETA, c, 1, n)
                flops_per_step = 2.0 * n ** 3
            else:
                m_t, n_t = 2 * n, n // 2
                g = torch.Generator(device=dev); g.manual_seed(0xB200)
                at = torch.rand(m_t, m_t, dtype=torch.float64, device=dev, generator=g) * 2 - 1
                at = at / float(at.abs().sum(dim=1).max()); at.diagonal().add_(2.0)
                at = at.t()                                       # column-major, lower triangle is what trsm reads
                bt0 = (torch.rand(n_t, m_t, dtype=torch.float64, device=dev, generator=g) * 2 - 1).t()
                bt = bt0.clone(memory_format=torch.preserve_format)

                def step():
                    bt.copy_(bt0)
                    api.bli_dtrsm(0, 0xC0, 0, 0, m_t, n_t, ALPHA, at, 1, m_t, bt, 1, m_t)
                flops_per_step = 1.0 * m_t * m_t * n_t
            total_flops = flops_per_step
            parallelism = "single GPU"
        elif args.op == "dtrsm":
            # multi-GPU trsm: B (and X) split into column blocks by bli_thread_range_sub, A replicated,
            # no data-path collective (the reference's jc/jr parallelism, frame/3/trsm/bli_trsm_cntl.c:446-451)
            from blis_b200 import dist as bdist
            m_t, n_t = 2 * n, n // 2
            j0, j1 = bdist.trsm_column_block(rank, world, n_t)
            g = torch.Generator(device=dev); g.manual_seed(0xB200)
            at = torch.rand(m_t, m_t, dtype=torch.float64, device=dev, generator=g) * 2 - 1
            at = at / float(at.abs().sum(dim=1).max()); at.diagonal().add_(2.0)
            at = at.t()
            bt0 = (torch.rand(j1 - j0, m_t, dtype=torch.float64, device=dev, generator=g) * 2 - 1).t()
            bt = bt0.clone(memory_format=torch.preserve_format)

            def step():
                bt.copy_(bt0)
                api.bli_dtrsm(0, 0xC0, 0, 0, m_t, j1 - j0, ALPHA, at, 1, m_t, bt, 1, m_t)
            flops_per_step = 1.0 * m_t * m_t * (j1 - j0)
            total_flops = 1.0 * m_t * m_t * n_t
            parallelism = f"B split into {world} column blocks (bli_thread_range_sub, bf=128), A replicated; strong scaling, no collective"
        else:
            from blis_b200 import dist as bdist
            kb = int(os.environ.get("B200_DIST_KB", "2048"))
            if args.workload == "g3":
                job = bdist.DistGemm(65536, 65536, 65536, world, rank, dev, alpha=ALPHA, beta=BETA, kb=kb)
                workload = "dgemm m=n=k=65536 column-major fp64, 2D-sharded, alpha=2.0 beta=1.2 (BASELINE configs[4])"
            else:
                job = bdist.WeakScalingGemm(n, world, rank, dev, alpha=ALPHA, beta=BETA, kb=kb)

            def step():
                job.step()
            total_flops = job.total_flops
            parallelism = job.describe()

        for _ in range(args.warmup):
            step()
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches[0] = api.launch_count()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        barrier()
        ev[0].record(stream)
        for i in range(args.steps):
            step()
            ev[i + 1].record(stream)
        barrier()
        launches[0] = api.launch_count() - launches[0]
        clocks = sampler.stop() if rank == 0 else None
        elapsed_ms = ev[0].elapsed_time(ev[-1])
        per_step = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]

        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        value = total_flops * args.steps / (elapsed_ms * 1e-3) / 1e9

        # ---- roofline of the dominant kernel (the dgemm tile kernel), timed live with CUDA events
        roof = None
        if rank == 0:
            # TFLOP/s, FP64 tensor pipe, this GPU, this run (profiling runs reuse the committed figure)
            peak = 37.08 if args.no_peak else api.measure_peak("dmma", 300)
            kern_ms = sum(per_step) / len(per_step)
            if world == 1 and args.op == "dgemm":
                achieved = flops_per_step / (kern_ms * 1e-3) / 1e12
            else:
                achieved = (total_flops / world) / (kern_ms * 1e-3) / 1e12
            traffic = None
            tfile = ROOT / "profiles" / "dgemm_traffic.json"
            if tfile.exists():
                try:
                    traffic = json.loads(tfile.read_text()).get("dram_bytes_per_launch")
                except (ValueError, OSError):
                    traffic = None
            roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic,
                    "peak_source": "FP64 DMMA (mma.sync m8n8k4) microbenchmark measured in this run by b200_measure_peak; "
                                   "MEASURED_PEAKS.json holds only HBM GB/s and bf16 TFLOP/s",
                    "kernel": "gemm_dmma_tma_kernel<XK=1,YK=0> (128x128x16, 6 stages, TMA)", "algorithmic_flop_per_launch": flops_per_step if world == 1 else total_flops / world}

        # ---- e2e: same call through the public API with HOST (pinned) operands
        e2e = None
        if not args.no_e2e and world == 1 and args.op == "dgemm":
            ah = torch.empty(n, n, dtype=torch.float64).pin_memory().t()
            bh = torch.empty(n, n, dtype=torch.float64).pin_memory().t()
            ch = torch.empty(n, n, dtype=torch.float64).pin_memory().t()
            ah.copy_(a); bh.copy_(b); ch.copy_(c)
            torch.cuda.synchronize()
            e2e_steps = max(2, min(args.steps, 3))
            api.bli_dgemm(0, 0, n, n, n, ALPHA, ah, 1, n, bh, 1, n, BETA, ch, 1, n)      # warm-up (allocations)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                api.bli_dgemm(0, 0, n, n, n, ALPHA, ah, 1, n, bh, 1, n, BETA, ch, 1, n)  # returns after C is back on the host
            dt = (time.perf_counter() - t0) / e2e_steps
            e2e = {"value": 2.0 * n ** 3 / dt / 1e9, "unit": "GFLOPS", "h2d_bytes_per_step": 3 * n * n * 8,
                   "d2h_bytes_per_step": n * n * 8, "ms_per_step": dt * 1e3,
                   "how": "bli_dgemm on pinned host operands: H2D of A,B,C + kernel + D2H of C per step, host wall clock"}

        # ---- e2e at N > 1: every rank keeps its shards (block-cyclic k panels of A and B, its block of C) in pinned HOST
        # memory; a step is DistGemm.step_host: shards uploaded in the order the k steps use them, C in column blocks under
        # the first k step, finished column blocks of C read back under the last one (blis_b200/dist.py: summa_host).
        # All ranks agree first that their pinned buffers exist, so that no rank can leave the others inside a collective.
        if not args.no_e2e and world > 1 and args.op == "dgemm" and args.workload == "headline":
            try:
                hosts, ok = job.host_shards(), 1
            except Exception as exc:                                     # noqa: BLE001 (pinning can fail on a small host)
                print(f"[rank {rank}] e2e skipped: {exc}", file=sys.stderr)
                hosts, ok = None, 0
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
              try:
                h2d = sum(h.numel() * h.element_size() for h in hosts)
                d2h = hosts[2].numel() * hosts[2].element_size()

                def e2e_step():
                    job.step_host(hosts)
                    torch.cuda.synchronize()                             # C is back in host memory
                e2e_steps = max(2, min(args.steps, 3))
                e2e_step()                                               # warm-up
                barrier()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    e2e_step()
                barrier()
                tt = torch.tensor([(time.perf_counter() - t0) / e2e_steps, float(h2d), float(d2h)], dtype=torch.float64, device=dev)
                tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                dist.all_reduce(tt, op=dist.ReduceOp.SUM)
                dt = float(tmax[0].item())
                e2e = {"value": total_flops / dt / 1e9, "unit": "GFLOPS", "h2d_bytes_per_step": int(tt[1].item()),
                       "d2h_bytes_per_step": int(tt[2].item()), "ms_per_step": dt * 1e3,
                       "how": "DistGemm.step_host: every rank's shards of A, B and C live in pinned host memory; H2D in k-step order, "
                              "C in column blocks under the first k step, C blocks read back under the last; host wall clock "
                              "between barriers, max over ranks; bytes summed over ranks"}
              except Exception as exc:                                   # noqa: BLE001 -- keep the device-resident line
                print(f"[rank {rank}] e2e failed: {exc!r}", file=sys.stderr)
                e2e = None

    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        v, cores, sample, _ = cpu_reference_gflops(args.op, budget_s=20.0, reps=2)
        cpu = {"value": v, "unit": "GFLOPS", "cores": cores, "kind": "reference", "sample": sample}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "GFLOPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if (args.op == "dtrsm" or args.workload == "g3") and world > 1 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "parallelism": parallelism,
                       "l2": "inputs (3 x 2 GiB per GPU) far exceed the 126 MB L2; no flush needed",
                       "timing": "CUDA events on the launching stream, max over ranks"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches[0], "roofline": roof, "cpu_baseline": cpu,
        }))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
