"""Multi-GPU path on CPU: world_size-2 gloo run of the 2D-decomposed gemm
pipeline (plan + all-gather exchange + panel loop) with the compute step
replaced by a CPU matmul, plus pure index checks of larger grids."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from blis_b200 import partition
from blis_b200.dist import (PanelExchange, SummaPlan, _NoStream, col_blocks, summa, summa_host, trsm_column_block,
                            trsm_host_blocks)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _global(M, N, K):
    g = torch.Generator(); g.manual_seed(1234)
    A = torch.rand(M, K, dtype=torch.float64, generator=g) * 2 - 1
    B = torch.rand(K, N, dtype=torch.float64, generator=g) * 2 - 1
    C = torch.rand(M, N, dtype=torch.float64, generator=g) * 2 - 1
    return A, B, C


def _worker(rank, world, port, M, N, K, kb, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = SummaPlan(world, rank, M, N, K, kb)
        A, B, C = _global(M, N, K)
        Ai, Bj = A[p.m0:p.m1], B[:, p.n0:p.n1]
        a_loc = torch.stack([Ai[:, t * kb:(t + 1) * kb].t().contiguous() for t in p.a_panels()])      # [na, kb, m_loc]
        b_loc = torch.stack([Bj[t * kb:(t + 1) * kb].t().contiguous() for t in p.b_panels()])        # [nb, n_loc, kb]
        c = C[p.m0:p.m1, p.n0:p.n1].clone()
        ex = PanelExchange(p, a_loc, b_loc)
        seen = []

        def gemm_panel(first, a_t, b_t):
            nonlocal c
            c = (1.2 * c if first else c) + 2.0 * (a_t.t() @ b_t.t())
            seen.append(1)
        summa(p, ex, gemm_panel)
        want = 1.2 * C[p.m0:p.m1, p.n0:p.n1] + 2.0 * (Ai @ Bj)
        q.put((rank, float((c - want).abs().max()), len(seen), (p.pr, p.pc)))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e), -1, None))
    finally:
        dist.destroy_process_group()


def _worker_host(rank, world, port, M, N, K, kb, q):
    """summa_host: shards start in 'host' tensors, the device shards are poisoned, C must come home complete."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = SummaPlan(world, rank, M, N, K, kb)
        A, B, C = _global(M, N, K)
        Ai, Bj = A[p.m0:p.m1], B[:, p.n0:p.n1]
        a_h = torch.stack([Ai[:, t * kb:(t + 1) * kb].t().contiguous() for t in p.a_panels()])
        b_h = torch.stack([Bj[t * kb:(t + 1) * kb].t().contiguous() for t in p.b_panels()])
        c_h = C[p.m0:p.m1, p.n0:p.n1].t().contiguous()                                  # dense [n_loc, m_loc]
        a_loc, b_loc, c_dense = (torch.full_like(x, float("nan")) for x in (a_h, b_h, c_h))
        ex = PanelExchange(p, a_loc, b_loc)
        calls = []

        def gemm_cols(first, a_ts, b_ts, j0, j1):
            acc = sum(b_t[j0:j1] @ a_t for a_t, b_t in zip(a_ts, b_ts))                    # (A_t B_t)^T restricted to the block
            c_dense[j0:j1] = (1.2 * c_dense[j0:j1] if first else c_dense[j0:j1]) + 2.0 * acc
            calls.append((first, j0, j1))
        ns = _NoStream()
        for _ in range(2):                                                               # twice: C is re-read from the host image
            c_h.copy_(C[p.m0:p.m1, p.n0:p.n1].t())
            summa_host(p, ex, (a_h, b_h, c_h), c_dense, gemm_cols, ns, ns, ns, nblk=3, bf=2)
        want = 1.2 * C[p.m0:p.m1, p.n0:p.n1] + 2.0 * (Ai @ Bj)
        q.put((rank, float((c_h.t() - want).abs().max()), calls, p.steps, col_blocks(p.n_loc, 3, 2)))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e), None, None, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("K", [64, 32, 16])      # 4 k steps; 2 (first, last: the N=8 bench shape); 1 (first == last)
def test_summa_host_shards_two_ranks_gloo(K):
    world, M, N, kb = 2, 48, 20, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_host, args=(r, world, port, M, N, K, kb, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    for rank, err, calls, steps, blocks in res:
        assert not isinstance(err, str), err
        assert err < 1e-12, (rank, err)
        assert len(blocks) > 1 and blocks[0][0] == 0
        per = calls[:len(calls) // 2]
        # first and last k step run block by block, the others as one launch over all columns
        want_calls = [(True, *b) for b in blocks]
        if steps > 1:
            want_calls += [(False, 0, blocks[-1][1])] * (steps - 2) + [(False, *b) for b in blocks]
        assert per == want_calls, (per, want_calls)


def test_summa_two_ranks_gloo():
    world, M, N, K, kb = 2, 48, 20, 64, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, M, N, K, kb, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    for rank, err, npanels, grid in res:
        assert not isinstance(err, str), err
        assert err < 1e-12, (rank, err)
        assert npanels == K // kb
        assert grid == partition.thread_partition_2x2(2, M, N)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_plan_covers_everything(world):
    n, kb = 16384, 1024
    pr, pc = partition.thread_partition_2x2(world, n, n)
    plans = [SummaPlan(world, r, pr * n, pc * n, n, kb) for r in range(world)]
    # C blocks tile the global matrix exactly
    cover = np.zeros((pr, pc), dtype=int)
    for p in plans:
        cover[p.i, p.j] += 1
        assert (p.m_loc, p.n_loc) == (n, n)
    assert (cover == 1).all()
    for p in plans:
        # every k panel is fetched exactly once per step sequence, from its owner
        ts = [t for s in range(p.steps) for (t, *_rest) in p.step_panels(s)]
        assert ts == list(range(p.T))
        for s in range(p.steps):
            for t, ja, qa, ib, qb in p.step_panels(s):
                owner_a = plans[p.i * p.pc + ja]; owner_b = plans[ib * p.pc + p.j]
                assert owner_a.a_panels()[s * (p.L // p.pc) + qa] == t
                assert owner_b.b_panels()[s * (p.L // p.pr) + qb] == t


def test_trsm_column_blocks_follow_thread_range():
    n = 8192
    for world in (1, 2, 4, 8):
        blocks = [trsm_column_block(r, world, n) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        assert blocks == [partition.thread_range_sub(r, world, n, 128) for r in range(world)]


def test_trsm_host_blocks_order_and_result():
    """trsm_host_blocks on CPU tensors: every column sub-block is uploaded, solved once, and brought home."""
    m, n_loc = 24, 20
    g = torch.Generator(); g.manual_seed(7)
    a = torch.tril(torch.rand(m, m, dtype=torch.float64, generator=g)) + 2.0 * torch.eye(m, dtype=torch.float64)
    b_host = torch.rand(n_loc, m, dtype=torch.float64, generator=g)               # dense image: row j = column j of B
    want = torch.linalg.solve_triangular(a, 2.0 * b_host.t(), upper=False).t()
    b_dev = torch.full_like(b_host, float("nan"))
    blocks = col_blocks(n_loc, 3, 4)
    assert len(blocks) == 3 and blocks[0][0] == 0 and blocks[-1][1] == n_loc
    calls = []

    def solve_cols(j0, j1):
        b_dev[j0:j1] = torch.linalg.solve_triangular(a, 2.0 * b_dev[j0:j1].t(), upper=False).t()
        calls.append((j0, j1))
    ns = _NoStream()
    trsm_host_blocks(solve_cols, b_dev, b_host, blocks, ns, ns, ns)
    assert calls == blocks
    assert float((b_host - want).abs().max()) < 1e-12
