"""Multi-GPU gemm/trsm: one process per GPU, torch.distributed (NCCL) plumbing.

The reference has no distributed layer ("Distributed
communication backend: none"); what it has is the jc x ic thread partitioning of
C (frame/base/bli_rntm.c:424-489 -> bli_thread_partition_2x2) and contiguous
per-thread ranges (bli_thread_range_sub).  This module applies exactly that
arithmetic (blis_b200.partition) across GPUs:

gemm  -- C is split into a Pr x Pc grid of blocks (Pr, Pc from
         thread_partition_2x2(world, M, N)); rank (i, j) owns C_ij.  k is never
         split across GPUs (as the reference never splits k across threads), so
         there is no reduction and a block's bits do not depend on the GPU count
         beyond the k-panel order.  A's row panel i is distributed over the Pc
         ranks of grid row i and B's column panel j over the Pr ranks of grid
         column j, both block-cyclically along k with panel width kb.  Each step
         all-gathers the next L = lcm(Pr, Pc) k-panels inside the row group (A)
         and the column group (B) -- NCCL all-gather over NVLink -- double
         buffered, so the gather of step s+1 runs under the DMMA kernels of step s.
trsm  -- B (and X) split into column blocks by thread_range_sub, A replicated:
         no collective on the data path, exactly the reference's jc/jr
         parallelism (frame/3/trsm/bli_trsm_cntl.c:446-451).

The index logic (`SummaPlan`) and the exchange (`PanelExchange`) are backend
agnostic so that they are tested on CPU with gloo (tests/test_dist_cpu.py).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import partition


@dataclass
class SummaPlan:
    """Pure index arithmetic of the 2D decomposition."""
    world: int
    rank: int
    M: int
    N: int
    K: int
    kb: int

    def __post_init__(self):
        self.pr, self.pc = partition.thread_partition_2x2(self.world, self.M, self.N)
        assert self.pr * self.pc == self.world
        self.i, self.j = divmod(self.rank, self.pc)
        self.L = math.lcm(self.pr, self.pc)
        self.T = -(-self.K // self.kb)                       # number of k panels
        if self.K % self.kb or self.T % self.L:
            raise ValueError(f"K={self.K} must be a multiple of kb*lcm(Pr,Pc)={self.kb * self.L}")
        self.steps = self.T // self.L
        # my block of C (ragged edge on the last rank, as bli_thread_range_sub does)
        self.m0, self.m1 = partition.thread_range_sub(self.i, self.pr, self.M, 1)
        self.n0, self.n1 = partition.thread_range_sub(self.j, self.pc, self.N, 1)

    @property
    def m_loc(self): return self.m1 - self.m0
    @property
    def n_loc(self): return self.n1 - self.n0

    def row_group(self):  return [self.i * self.pc + jj for jj in range(self.pc)]
    def col_group(self):  return [ii * self.pc + self.j for ii in range(self.pr)]
    def a_panels(self):   return [t for t in range(self.T) if t % self.pc == self.j]      # k panels of A I own
    def b_panels(self):   return [t for t in range(self.T) if t % self.pr == self.i]      # k panels of B I own

    def step_panels(self, s):
        """For step s: list of (t, a_src_rank_in_row, a_slot, b_src_rank_in_col, b_slot)."""
        out = []
        for t in range(s * self.L, (s + 1) * self.L):
            out.append((t, t % self.pc, (t - s * self.L) // self.pc, t % self.pr, (t - s * self.L) // self.pr))
        return out


class PanelExchange:
    """All-gather of the step's A panels in the row group and B panels in the column group."""

    def __init__(self, plan: SummaPlan, a_loc: torch.Tensor, b_loc: torch.Tensor):
        # a_loc: [n_a_panels, kb, m_loc]  (panel, k, m)  -> each panel is column-major m_loc x kb
        # b_loc: [n_b_panels, n_loc, kb]  (panel, n, k)  -> each panel is column-major kb x n_loc
        self.p, self.a_loc, self.b_loc = plan, a_loc, b_loc
        self.row_pg = self.col_pg = None
        # every rank must create every group, in the same order
        for i in range(plan.pr):
            g = dist.new_group([i * plan.pc + jj for jj in range(plan.pc)])
            if i == plan.i:
                self.row_pg = g
        for j in range(plan.pc):
            g = dist.new_group([ii * plan.pc + j for ii in range(plan.pr)])
            if j == plan.j:
                self.col_pg = g
        self.qa, self.qb = plan.L // plan.pc, plan.L // plan.pr       # panels I contribute per step
        kw = dict(dtype=a_loc.dtype, device=a_loc.device)
        self.abuf = [torch.empty(plan.pc, self.qa, plan.kb, plan.m_loc, **kw) for _ in range(2)]
        self.bbuf = [torch.empty(plan.pr, self.qb, plan.n_loc, plan.kb, **kw) for _ in range(2)]

    def start(self, s: int):
        """Launch the (asynchronous) gathers of step s into buffer s%2; returns work handles."""
        p = self.p
        a_src = self.a_loc[s * self.qa:(s + 1) * self.qa]
        b_src = self.b_loc[s * self.qb:(s + 1) * self.qb]
        # flat views: rank r's contribution lands in slot [r] of the (contiguous) buffer
        wa = dist.all_gather_into_tensor(self.abuf[s % 2].view(-1), a_src.reshape(-1), group=self.row_pg, async_op=True)
        wb = dist.all_gather_into_tensor(self.bbuf[s % 2].view(-1), b_src.reshape(-1), group=self.col_pg, async_op=True)
        return wa, wb

    def panels(self, s: int):
        """After the step's gathers completed: yields (t, A_t [kb, m_loc], B_t [n_loc, kb])."""
        ab, bb = self.abuf[s % 2], self.bbuf[s % 2]
        for t, ja, qa, ib, qb in self.p.step_panels(s):
            yield t, ab[ja, qa], bb[ib, qb]


def summa(plan: SummaPlan, ex: PanelExchange, gemm_panel, n_steps=None, gemm_step=None):
    """Drive the pipeline: gemm_panel(first, A_t, B_t) accumulates one k panel into C_ij, or, when
    gemm_step is given, gemm_step(first, [A_t...], [B_t...]) accumulates all panels of a step at once."""
    steps = plan.steps if n_steps is None else n_steps
    works = {0: ex.start(0)}
    if steps > 1:
        works[1] = ex.start(1)
    first = True
    trace = [] if os.environ.get("B200_DIST_TRACE") == "1" else None
    for s in range(steps):
        if trace is not None:
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            trace.append(e); e[0].record()
        for w in works.pop(s):
            w.wait()
        if trace is not None:
            trace[-1][1].record()
        if gemm_step is not None:
            ps = list(ex.panels(s))
            gemm_step(first, [p[1] for p in ps], [p[2] for p in ps])
            first = False
        else:
            for _, a_t, b_t in ex.panels(s):
                gemm_panel(first, a_t, b_t)
                first = False
        if trace is not None:
            trace[-1][2].record()
        if s + 2 < steps:
            works[s + 2] = ex.start(s + 2)       # ordered after this step's kernels: its buffer is free again
    if trace is not None:
        torch.cuda.synchronize()
        print(f"[rank {plan.rank}] step (wait_ms, gemm_ms): " +
              " ".join(f"({e[0].elapsed_time(e[1]):.1f},{e[1].elapsed_time(e[2]):.1f})" for e in trace), flush=True)


class _NoStream:
    """Stand-in for CUDA streams/events on CPU tensors (tests/test_dist_cpu.py): everything is already ordered."""
    def wait_stream(self, other): pass
    def wait_event(self, ev): pass
    def record_event(self): return None
    def __enter__(self): return self
    def __exit__(self, *exc): return False


class _CudaStream:
    """The three operations summa_host needs from a stream, on a torch.cuda stream."""
    def __init__(self, stream): self.s = stream
    def wait_stream(self, other): self.s.wait_stream(other.s)
    def wait_event(self, ev): self.s.wait_event(ev)

    def record_event(self):
        ev = torch.cuda.Event(); ev.record(self.s); return ev

    def __enter__(self):
        self._ctx = torch.cuda.stream(self.s); self._ctx.__enter__(); return self

    def __exit__(self, *exc): return self._ctx.__exit__(*exc)


def col_blocks(n_loc: int, nblk: int, bf: int = 128):
    """Column blocks of the local C block, split as bli_thread_range_sub does (multiples of bf, edge on the last)."""
    out = [partition.thread_range_sub(j, nblk, n_loc, bf) for j in range(nblk)]
    return [(j0, j1) for j0, j1 in out if j1 > j0]


def summa_host(plan: SummaPlan, ex: PanelExchange, hosts, c_dense: torch.Tensor, gemm_cols, cur, s_in, s_out, nblk: int = 4,
               bf: int = 128, s_nc=None):
    """One distributed product whose shards live in (pinned) HOST memory: the host-pointer path of the multi-GPU gemm.

    hosts = (a_h, b_h, c_h): host images of ex.a_loc [na, kb, m_loc], ex.b_loc [nb, n_loc, kb] and of the dense
    [n_loc, m_loc] tensor behind the rank's column-major C block (row j of it = column j of C).
    gemm_cols(first, a_ts, b_ts, j0, j1) accumulates the step's panels into columns [j0, j1) of C.
    Pipeline (cur = compute stream, s_in = H2D stream, s_out = D2H stream, s_nc = the stream the all-gathers are issued
    from; cur if None):
      * s_in uploads the shards in the order the k steps use them; the all-gather of step s is ordered behind upload s
        -- on s_nc, not on cur: a wait for upload s+1 queued on the compute stream ahead of step s's kernels would hold
        them back until the whole of C (which travels between the two uploads) has arrived (measured: 70 ms);
      * the host C arrives in column blocks right behind the first step's shards, and the FIRST k step runs block by
        block, each launch waiting only for its block of C;
      * the LAST k step runs block by block as well, and every finished block goes home on s_out under the next one.
    Exposed transfer: the first step's shards + one block of C in, one block of C out."""
    a_h, b_h, c_h = hosts
    p, steps = plan, plan.steps
    blocks = col_blocks(p.n_loc, nblk, bf)
    s_in.wait_stream(cur)                       # the previous product has finished with the device shards
    s_out.wait_stream(cur)
    ev_up, ev_c = [], []
    with s_in:
        for s in range(steps):
            ex.a_loc[s * ex.qa:(s + 1) * ex.qa].copy_(a_h[s * ex.qa:(s + 1) * ex.qa], non_blocking=True)
            ex.b_loc[s * ex.qb:(s + 1) * ex.qb].copy_(b_h[s * ex.qb:(s + 1) * ex.qb], non_blocking=True)
            ev_up.append(s_in.record_event())
            if s == 0:
                for j0, j1 in blocks:
                    c_dense[j0:j1].copy_(c_h[j0:j1], non_blocking=True)
                    ev_c.append(s_in.record_event())
    works = {}
    if s_nc is None:
        s_nc = cur
    else:
        s_nc.wait_stream(cur)

    def start(s, buffer_free=None):
        with s_nc:
            s_nc.wait_event(ev_up[s])
            if buffer_free is not None:
                s_nc.wait_event(buffer_free)     # the kernels that read gather buffer s % 2 have finished
            works[s] = ex.start(s)
    start(0)
    if steps > 1:
        start(1)
    for s in range(steps):
        for w in works.pop(s):
            w.wait()
        ps = list(ex.panels(s))
        a_ts, b_ts = [q[1] for q in ps], [q[2] for q in ps]
        last = s == steps - 1
        for jb, (j0, j1) in enumerate(blocks if (s == 0 or last) else [(0, p.n_loc)]):
            if s == 0:
                cur.wait_event(ev_c[jb])
            gemm_cols(s == 0, a_ts, b_ts, j0, j1)
            if last:
                s_out.wait_event(cur.record_event())
                with s_out:
                    c_h[j0:j1].copy_(c_dense[j0:j1], non_blocking=True)
        if s + 2 < steps:
            start(s + 2, cur.record_event())
    cur.wait_stream(s_out)                      # whoever synchronises `cur` has C at home


_native_up = False


def native_init(device) -> None:
    """Bring up the engine's own NCCL communicator (b200_dist_init) once per process, bootstrapped over the already
    initialised torch.distributed group (api.dist_init)."""
    global _native_up
    if not _native_up:
        from . import api
        api.dist_init(device)
        _native_up = True


def native_finalize() -> None:
    global _native_up
    if _native_up:
        from . import api
        torch.cuda.synchronize()
        api.dist_finalize()
        _native_up = False


class DistGemm:
    """C := beta*C + alpha*A*B on a Pr x Pc grid of GPUs; every rank holds its block of C and its
    block-cyclic k-panels of A's row panel / B's column panel (synthetic data generated in place).

    `step()` is ONE call of the C ABI, b200_dist_gemm (blis_b200/csrc/host_dist.cuh): the whole pipeline -- gathers on the
    engine's communication stream, double buffering, one k-panel launch per step -- runs in the engine.  The Python
    pipeline below (`summa`, `PanelExchange`) is the same schedule written against torch.distributed; it serves the gloo
    test of the index logic on CPU, the host-shard path (`step_host`) and, in `verify()`, as the independent replay."""

    def __init__(self, M: int, N: int, K: int, world: int, rank: int, device, alpha=2.0, beta=1.2, kb: int = 2048, native: bool = True):
        from . import api
        self.api = api
        self.native = bool(native) and torch.device(device).type == "cuda"
        if self.native:
            native_init(device)
        self.plan = SummaPlan(world, rank, M, N, K, kb)
        p = self.plan
        self.alpha, self.beta = alpha, beta
        g = torch.Generator(device=device); g.manual_seed(0xB200 + rank)
        scale = 1.0 / K
        na, nb = len(p.a_panels()), len(p.b_panels())
        self.a_loc = (torch.rand(na, kb, p.m_loc, dtype=torch.float64, device=device, generator=g) * 2 - 1) * scale
        self.b_loc = (torch.rand(nb, p.n_loc, kb, dtype=torch.float64, device=device, generator=g) * 2 - 1) * scale
        self.c = (torch.rand(p.n_loc, p.m_loc, dtype=torch.float64, device=device, generator=g) * 2 - 1).t()
        self.ex = PanelExchange(p, self.a_loc, self.b_loc)
        self.total_flops = 2.0 * p.M * p.N * p.K
        self.launches_per_step = p.steps
        # static shards: share them between the ranks once, so that the k panels are pulled by the copy engines
        self.one_sided = False
        if self.native and world > 1 and os.environ.get("B200_DIST_TRANSPORT", "ce") != "nccl":
            self.one_sided = self.api.dist_register(self.a_loc, self.b_loc)

    def close(self):
        if getattr(self, "one_sided", False):
            self.api.dist_unregister(self.a_loc, self.b_loc)
            self.one_sided = False

    def describe(self) -> str:
        p = self.plan
        return (f"2D block decomposition of C on a {p.pr}x{p.pc} grid (bli_thread_partition_2x2), C_ij {p.m_loc}x{p.n_loc} per GPU, "
                f"global {p.M}x{p.N}x{p.K}; A/B k-panels (kb={p.kb}) gathered in row/column groups "
                + ("by one-sided peer copies (copy engines over NVLink, registered static shards), " if self.one_sided else "with NCCL all-gather, ")
                + f"double buffered under the DMMA kernels; no reduction (k not split across GPUs); "
                + ("one b200_dist_gemm call per product (pipeline inside the engine, C ABI)" if self.native else "pipeline driven from Python"))

    def _panel(self, first, a_t, b_t):
        p = self.plan
        # a_t: [kb, m_loc] contiguous == column-major m_loc x kb (rs=1, cs=m_loc); b_t: [n_loc, kb] == column-major kb x n_loc
        self.api.bli_dgemm(0, 0, p.m_loc, p.n_loc, p.kb, self.alpha, a_t, 1, p.m_loc, b_t, 1, p.kb,
                           self.beta if first else 1.0, self.c, 1, p.m_loc)

    def _step_panels(self, first, a_ts, b_ts):
        """All L panels of a step in ONE launch (b200_gemm_kpanels): the k loop runs over L*kb."""
        p = self.plan
        for lo in range(0, len(a_ts), 8):
            self.api.bli_gemm_kpanels(torch.float64, 0, 0, p.m_loc, p.n_loc, p.kb, self.alpha, a_ts[lo:lo + 8], 1, p.m_loc,
                                      b_ts[lo:lo + 8], 1, p.kb, self.beta if (first and lo == 0) else 1.0, self.c, 1, p.m_loc)

    def step(self, flags: int | None = None):
        p = self.plan
        if not self.native:
            return summa(p, self.ex, self._panel, gemm_step=self._step_panels)
        # A and B never change between the products of this job: their first gather may run under the previous product
        self.api.dist_gemm(torch.float64, p.M, p.N, p.K, p.kb, self.alpha, self.a_loc, self.b_loc, self.beta, self.c, 1, p.m_loc,
                           self.api.DIST_AB_STATIC if flags is None else flags)

    def step_py(self):
        """The same product with the pipeline driven from Python over torch.distributed (what step() was in round 1)."""
        summa(self.plan, self.ex, self._panel, gemm_step=self._step_panels)

    # ---- parity of the distributed result ----------------------------------------------------------------------------
    def verify(self, hosts=None) -> dict:
        """Post-timing check of ONE distributed product, run on every rank (collective):

        * bit_equal -- the rank's block of C after `step()` equals, bit for bit, the block the single-GPU engine computes
          when it replays the same k-panel schedule (same b200_gemm_kpanels calls, beta then 1) on panels gathered
          INDEPENDENTLY of PanelExchange (whole-shard all_gather in the row/column group, re-indexed by panel number).
          k is never split across GPUs (frame/3/gemm/bli_gemm_blk_var3.c:110-112: the reference never splits k across
          threads either), so a wrong slot mapping, a buffer-reuse race between the gather of step s+2 and the kernels of
          step s, or a panel-segment bug shows up as a bit difference.
        * resid -- the reference testsuite's randomized residual of that block, || C t - (beta C0 t + alpha A (B t)) ||
          (testsuite/src/test_gemm.c:393-401; pass threshold 1e-14 at :44-47), computed with torch mat-vecs.
        * host_bit_equal (when `hosts` is given) -- step_host() from the same C0 in pinned host memory gives the same bits."""
        p, ex = self.plan, self.ex
        cd = self.c.t()                                       # dense [n_loc, m_loc] tensor behind the column-major block
        c0 = cd.clone()
        self.step()
        torch.cuda.synchronize()
        c_dist = cd.clone()
        # whole shards into ONE tensor per operand: the replay's panels are then slots of one strided buffer, like the
        # engine's receive buffers, so both runs are served by the same kernel (its name is compared below)
        kern_dist = self.api.last_kernel()
        a_all = torch.empty((p.pc,) + tuple(self.a_loc.shape), dtype=self.a_loc.dtype, device=self.a_loc.device)
        b_all = torch.empty((p.pr,) + tuple(self.b_loc.shape), dtype=self.b_loc.dtype, device=self.b_loc.device)
        dist.all_gather_into_tensor(a_all.view(-1), self.a_loc.reshape(-1), group=ex.row_pg)
        dist.all_gather_into_tensor(b_all.view(-1), self.b_loc.reshape(-1), group=ex.col_pg)
        a_of = lambda t: a_all[t % p.pc][t // p.pc]           # noqa: E731  panel t of A: owner column t % Pc, local index t // Pc
        b_of = lambda t: b_all[t % p.pr][t // p.pr]           # noqa: E731
        cd.copy_(c0)
        for s in range(p.steps):
            ts = range(s * p.L, (s + 1) * p.L)
            self._step_panels(s == 0, [a_of(t) for t in ts], [b_of(t) for t in ts])
        torch.cuda.synchronize()
        bit_equal = bool(torch.equal(cd, c_dist))
        kern_replay = self.api.last_kernel()
        g = torch.Generator(device=cd.device); g.manual_seed(7 + p.rank)
        tv = (torch.rand(p.n_loc, dtype=cd.dtype, device=cd.device, generator=g) * 2 - 1) / p.N
        z = self.beta * (c0.t() @ tv)
        for t in range(p.T):
            z += self.alpha * (a_of(t).t() @ (b_of(t).t() @ tv))
        resid = float(torch.linalg.vector_norm(c_dist.t() @ tv - z))
        out = {"bit_equal": bit_equal, "resid": resid, "kernel": kern_dist, "kernel_replay": kern_replay,
               "how": "C_ij after step() vs the same b200_gemm_kpanels schedule replayed on whole-shard all_gather'ed panels (bit for bit), "
                      "and the testsuite residual ||C t - (beta C0 t + alpha A (B t))|| of the distributed block"}
        if hosts is not None:
            hosts[2].copy_(c0)
            self.step_host(hosts)
            torch.cuda.synchronize()
            out["host_bit_equal"] = bool(torch.equal(hosts[2].to(cd.device), c_dist))
        cd.copy_(c_dist)
        return out

    # ---- shards in pinned host memory ------------------------------------------------------------------------------
    def host_shards(self):
        """Pinned host images of this rank's shards (A panels, B panels, dense tensor behind the C block), filled from
        the device."""
        hosts = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in (self.a_loc, self.b_loc, self.c.t())]
        for h, d in zip(hosts, (self.a_loc, self.b_loc, self.c.t())):
            h.copy_(d)
        return hosts

    def _cols(self, first, a_ts, b_ts, j0, j1):
        """The step's panels accumulated into columns [j0, j1) of C_ij: one launch per 8 panels."""
        p = self.plan
        c_blk = self.c[:, j0:j1]                             # column-major m_loc x w view
        b_blk = [b[j0:j1] for b in b_ts]                     # [w, kb] rows of a panel == column-major kb x w
        for lo in range(0, len(a_ts), 8):
            self.api.bli_gemm_kpanels(torch.float64, 0, 0, p.m_loc, j1 - j0, p.kb, self.alpha, a_ts[lo:lo + 8], 1, p.m_loc,
                                      b_blk[lo:lo + 8], 1, p.kb, self.beta if (first and lo == 0) else 1.0, c_blk, 1, p.m_loc)

    def step_host(self, hosts, nblk: int = 4):
        """One product with the shards in pinned host memory (summa_host); returns with the work queued: synchronise the
        current stream (or the device) to have the C block back in hosts[2]."""
        if not hasattr(self, "_s_in"):
            self._s_in, self._s_out, self._s_nc = (torch.cuda.Stream(self.c.device) for _ in range(3))
        cur = _CudaStream(torch.cuda.current_stream(self.c.device))
        summa_host(self.plan, self.ex, hosts, self.c.t(), self._cols, cur, _CudaStream(self._s_in), _CudaStream(self._s_out), nblk,
                   s_nc=_CudaStream(self._s_nc))


class WeakScalingGemm(DistGemm):
    """bench.py's N>1 workload: every rank owns an n x n block of C of the
    (Pr*n) x (Pc*n) x n product; per-GPU flops equal the single-GPU workload."""

    def __init__(self, n: int, world: int, rank: int, device, alpha=2.0, beta=1.2, kb: int = 2048):
        pr, pc = partition.thread_partition_2x2(world, n, n)     # grid for equal work per dimension
        super().__init__(pr * n, pc * n, n, world, rank, device, alpha, beta, kb)
        assert (self.plan.pr, self.plan.pc) == (pr, pc)


def trsm_host_blocks(solve_cols, b_dev: torch.Tensor, b_host: torch.Tensor, blocks, cur, s_in, s_out):
    """X := alpha * inv(op(A)) * B for a column block of B that lives in (pinned) HOST memory.

    Columns of B are independent in trsm -- the reference's only parallel trsm loops run over them
    (frame/3/trsm/bli_trsm_cntl.c:446-451, bli_trsm_ll_ker_var2.c:209-212) -- so the block is cut into column sub-blocks:
    sub-block j+1 travels to the device on s_in and sub-block j-1 travels home on s_out while sub-block j is solved on cur.
    b_dev / b_host: the dense [n_loc, m] tensors behind the column-major m x n_loc block (row j = column j of B);
    solve_cols(j0, j1) solves columns [j0, j1) in place on the device.  Exposed transfer: the first sub-block in, the last
    one out."""
    s_in.wait_stream(cur)
    s_out.wait_stream(cur)
    ev_in = []
    with s_in:
        for j0, j1 in blocks:
            b_dev[j0:j1].copy_(b_host[j0:j1], non_blocking=True)
            ev_in.append(s_in.record_event())
    for jb, (j0, j1) in enumerate(blocks):
        cur.wait_event(ev_in[jb])
        solve_cols(j0, j1)
        s_out.wait_event(cur.record_event())
        with s_out:
            b_host[j0:j1].copy_(b_dev[j0:j1], non_blocking=True)
    cur.wait_stream(s_out)


class DistTrsm:
    """Left-side trsm on `world` GPUs: B (and X) split into column blocks by bli_thread_range_sub, the triangular A
    replicated, no data-path collective.  Every rank holds A and its block of B (synthetic data generated in place)."""

    def __init__(self, m: int, n: int, world: int, rank: int, device, alpha=2.0, uplo=0xC0, trans=0, diag=0, seed=0xB200, native: bool = True):
        from . import api
        self.api, self.m, self.n, self.alpha, self.uplo, self.trans, self.diag = api, m, n, alpha, uplo, trans, diag
        self.native = bool(native) and world > 1 and torch.device(device).type == "cuda"
        if self.native:
            native_init(device)
        self.j0, self.j1 = trsm_column_block(rank, world, n)
        self.n_loc = self.j1 - self.j0
        g = torch.Generator(device=device); g.manual_seed(seed)
        a = torch.rand(m, m, dtype=torch.float64, device=device, generator=g) * 2 - 1
        a = a / float(a.abs().sum(dim=1).max()); a.diagonal().add_(2.0)     # testsuite: random, then diag += 2
        self.a = a.t()                                                      # column-major m x m
        g.manual_seed(seed + 1 + rank)
        self.b0 = torch.rand(self.n_loc, m, dtype=torch.float64, device=device, generator=g) * 2 - 1   # dense image of B
        self.b = self.b0.clone()
        self.flops = 1.0 * m * m * self.n_loc
        self.total_flops = 1.0 * m * m * n

    def _solve_cols(self, j0, j1):
        bv = self.b[j0:j1].t()                                # column-major m x w view
        self.api.bli_dtrsm(0, self.uplo, self.trans, self.diag, self.m, j1 - j0, self.alpha, self.a, 1, self.m, bv, 1, self.m)

    def step(self, root: int = -1):
        """One solve of this rank's column block.  Native: ONE b200_dist_trsm call (the engine cuts the same block with
        bli_thread_range_sub; root >= 0 broadcasts A from that rank first, root < 0: A is already replicated)."""
        self.b.copy_(self.b0)
        if self.native:
            self.api.dist_trsm(torch.float64, 0, self.uplo, self.trans, self.diag, root, self.m, self.n, self.alpha,
                               self.a, 1, self.m, self.b.t(), 1, self.m)
        else:
            self._solve_cols(0, self.n_loc)

    def host_block(self):
        h = torch.empty(self.b0.shape, dtype=self.b0.dtype).pin_memory()
        h.copy_(self.b0)
        return h

    def step_host(self, b_host, nblk: int | None = None):
        """The rank's block of B lives in pinned host memory and X comes back to it (trsm_host_blocks); returns with the
        work queued.  The diagonal-block solves are latency bound and do not get faster for fewer columns, so every
        extra sub-block repeats their cost: measured on a B200 at m=8192, n=4096, four sub-blocks take 26.4 ms against
        21.0 ms for upload -> solve -> download in sequence (11.6 ms device-resident; profiles/r01_dist_trsm_host_check.json).
        Default: one block (sequential) unless the block is large enough for two halves to hide more transfer than the
        second pass over the diagonal costs."""
        if nblk is None:
            nblk = 2 if (self.m >= 16384 and self.n_loc >= 4096) else 1
        if not hasattr(self, "_s_in"):
            self._s_in, self._s_out = torch.cuda.Stream(self.b.device), torch.cuda.Stream(self.b.device)
        cur = _CudaStream(torch.cuda.current_stream(self.b.device))
        trsm_host_blocks(self._solve_cols, self.b, b_host, col_blocks(self.n_loc, nblk), cur, _CudaStream(self._s_in),
                         _CudaStream(self._s_out))


class DistSkinnyGemm:
    """Skinny product C (m x n) := beta*C + alpha*A (m x k) * B (k x n) with k << m, n (the shapes the reference's sup path
    serves, frame/3/bli_l3_sup.c:37-135) on `world` GPUs: 1-D split of C's COLUMNS in units of 128 (bli_thread_range_sub),
    every rank holds its columns of B and C, and the small operand A is broadcast once from rank `root`.
    ONE b200_dist_gemm_1d call per product."""

    def __init__(self, m: int, n: int, k: int, world: int, rank: int, device, alpha=2.0, beta=1.2, root: int = 0, dtype=torch.float64, seed=0xB200):
        from . import api
        self.api, self.m, self.n, self.k, self.alpha, self.beta, self.root, self.dtype = api, m, n, k, alpha, beta, root, dtype
        self.world, self.rank = world, rank
        if world > 1 and torch.device(device).type == "cuda":
            native_init(device)
        self.j0, self.j1 = partition.thread_range_sub(rank, world, n, 128)
        self.n_loc = self.j1 - self.j0
        g = torch.Generator(device=device); g.manual_seed(seed)
        rnd = lambda *shape: (torch.rand(*shape, dtype=torch.float64, device=device, generator=g) * 2 - 1).to(dtype)   # noqa: E731
        a_root = rnd(k, m) / k                                  # dense image of the column-major m x k operand
        # only `root` holds A before the product; the other ranks' buffers are poisoned so that the broadcast is needed
        self.a = a_root if (rank == root or root < 0) else torch.full_like(a_root, float("nan"))
        self.a_ref = a_root
        g.manual_seed(seed + 1 + rank)
        self.b = rnd(max(self.n_loc, 1), k)[: self.n_loc]       # dense image of column-major k x n_loc
        self.c0 = rnd(max(self.n_loc, 1), m)[: self.n_loc]      # dense image of column-major m x n_loc
        self.c = self.c0.clone()
        self.total_flops = 2.0 * m * n * k
        self.bytes_rank = self.c.element_size() * (m * k + k * self.n_loc + 2 * m * self.n_loc)

    def describe(self) -> str:
        return (f"1-D split of C's columns over {self.world} GPUs (bli_thread_range_sub, bf=128): C {self.m}x{self.n_loc} per GPU of {self.m}x{self.n}, "
                f"k={self.k}; A ({self.m}x{self.k}) broadcast from rank {self.root} per product (ncclBroadcast), B and C local; no reduction")

    def step(self):
        if self.world == 1:       # one GPU: the plain C-ABI gemm (no communicator exists)
            fn = {torch.float64: self.api.bli_dgemm, torch.float32: self.api.bli_sgemm,
                  torch.complex128: self.api.bli_zgemm, torch.complex64: self.api.bli_cgemm}[self.dtype]
            return fn(0, 0, self.m, self.n, self.k, self.alpha, self.a.t(), 1, self.m, self.b.t(), 1, self.k, self.beta, self.c.t(), 1, self.m)
        self.api.dist_gemm_1d(self.dtype, self.api.DIST_COLS, self.root, self.m, self.n, self.k, self.alpha,
                              self.a.t(), 1, self.m, self.b.t(), 1, self.k, self.beta, self.c.t(), 1, self.m)

    def verify(self) -> dict:
        """The rank's block against the single-GPU engine on the same operands (bit for bit: same kernel, same k order)
        and the testsuite residual (testsuite/src/test_gemm.c:393-401)."""
        self.c.copy_(self.c0)
        if self.world > 1 and self.rank != self.root and self.root >= 0:
            self.a.fill_(float("nan"))
        self.step()
        torch.cuda.synchronize()
        got = self.c.clone()
        a_ok = bool(torch.equal(self.a, self.a_ref))
        if self.n_loc == 0:
            return {"bit_equal": a_ok, "resid": 0.0}
        ref = self.c0.clone()
        fn = {torch.float64: self.api.bli_dgemm, torch.float32: self.api.bli_sgemm,
              torch.complex128: self.api.bli_zgemm, torch.complex64: self.api.bli_cgemm}[self.dtype]
        fn(0, 0, self.m, self.n_loc, self.k, self.alpha, self.a_ref.t(), 1, self.m, self.b.t(), 1, self.k, self.beta, ref.t(), 1, self.m)
        torch.cuda.synchronize()
        g = torch.Generator(device=got.device); g.manual_seed(7 + self.rank)
        tv = ((torch.rand(self.n_loc, dtype=torch.float64, device=got.device, generator=g) * 2 - 1) / self.n).to(self.dtype)
        z = self.beta * (self.c0.t() @ tv) + self.alpha * (self.a_ref.t() @ (self.b.t() @ tv))
        resid = float(torch.linalg.vector_norm(got.t() @ tv - z))
        self.c.copy_(got)
        return {"bit_equal": a_ok and bool(torch.equal(got, ref)), "resid": resid}


def trsm_column_block(rank: int, world: int, n: int, nr: int = 128):
    """Columns [start, end) of B that `rank` solves (multi-GPU trsm: B split by column blocks)."""
    return partition.thread_range_sub(rank, world, n, nr)
