"""The multi-GPU entry points of the C ABI (b200_dist_gemm / b200_dist_gemm_1d / b200_dist_trsm, blis_b200/csrc/host_dist.cuh)
on real shards, checked on every rank.  Run under torchrun at any world size (or stand-alone: world size 1, where the
engine's communicator has one rank); prints one JSON line per rank, every key ending in `_ok` must be true.

Checks (k is never split across GPUs -- frame/3/gemm/bli_gemm_blk_var3.c:110-112 -- so a block's bits depend only on the
k-panel order, and the single-GPU engine replaying that order must reproduce them EXACTLY):
  gemm    * bit-for-bit against the single-GPU engine replaying the k-panel schedule on independently gathered shards
            (DistGemm.verify), same kernel name on both sides, testsuite residual (testsuite/src/test_gemm.c:393-401)
          * against ONE plain b200_gemm of the whole product on gathered A, B, C (every rank computes it; tolerance)
          * ragged m, n (last rank's block smaller), several k steps, 1 step, z
          * three products back to back with B200_DIST_AB_STATIC (gather of the next product under the current one) and a
            product whose A shard is written on the stream right before the call (no static flag)
  skinny  * b200_dist_gemm_1d, columns and rows, d and s: bit-for-bit against the single-GPU engine on the rank's slice,
            the broadcast operand poisoned on every rank but the root beforehand
  trsm    * b200_dist_trsm with A broadcast from rank 0 (poisoned elsewhere): bit-for-bit against b200_trsm on the rank's
            column block, testsuite residual form (test_trsm.c:362-381)
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from blis_b200 import api, partition             # noqa: E402
from blis_b200 import dist as bdist              # noqa: E402

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29543")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
out = {"rank": rank, "world": world}


def full_product_check(job):
    """Gather the whole A, B, C0 on every rank, run ONE single-GPU b200_gemm, compare my block (tolerance: another kernel /
    another k order may serve the full-size call)."""
    p = job.plan
    c0 = job.c.clone(memory_format=torch.preserve_format)
    # my row block of A as a dense [K, m_loc] image from the row group's shards; likewise B
    a_all = torch.empty((p.pc,) + tuple(job.a_loc.shape), dtype=job.a_loc.dtype, device=dev)
    b_all = torch.empty((p.pr,) + tuple(job.b_loc.shape), dtype=job.b_loc.dtype, device=dev)
    dist.all_gather_into_tensor(a_all.view(-1), job.a_loc.reshape(-1), group=job.ex.row_pg)
    dist.all_gather_into_tensor(b_all.view(-1), job.b_loc.reshape(-1), group=job.ex.col_pg)
    a_rows = torch.cat([a_all[t % p.pc][t // p.pc] for t in range(p.T)], dim=0)          # [K, m_loc]   (column-major m_loc x K)
    b_cols = torch.cat([b_all[t % p.pr][t // p.pr] for t in range(p.T)], dim=1)          # [n_loc, K]   (column-major K x n_loc)
    want = c0.clone(memory_format=torch.preserve_format)
    api.bli_dgemm(0, 0, p.m_loc, p.n_loc, p.K, job.alpha, a_rows, 1, p.m_loc, b_cols, 1, p.K, job.beta, want, 1, p.m_loc)
    job.c.copy_(c0)
    job.step(); torch.cuda.synchronize()
    err = float((job.c - want).abs().max())
    job.c.copy_(c0)
    return err


try:
    pr, pc = partition.thread_partition_2x2(world, 1000, 1000)
    import math
    L = math.lcm(pr, pc)
    # ---------------------------------------------------------------- gemm on the grid
    for tag, (M, N, K, kb) in {"4steps": (pr * 1536, pc * 1536, 4 * 512 * L, 512),
                               "1step": (pr * 1024, pc * 1024, 256 * L, 256),
                               "ragged": (pr * 1536 + 37, pc * 1280 + 5, 2 * 384 * L, 384)}.items():
        job = bdist.DistGemm(M, N, K, world, rank, dev, alpha=2.0, beta=1.2, kb=kb)
        assert job.native
        c_init = job.c.clone(memory_format=torch.preserve_format)
        ck = job.verify()
        out[f"gemm_{tag}_transport"] = api.dist_transport()
        if world > 1:
            # the one-sided transport (registered shards pulled by the copy engines) served it -- or, where the box cannot share
            # memory between processes, registration failed on every rank alike and NCCL did
            out[f"gemm_{tag}_transport_ok"] = (api.dist_transport() == ("copy-engine gets" if job.one_sided else "NCCL all-gather"))
            if tag == "4steps" and job.one_sided:
                # the same product over NCCL (registration dropped): same bits
                ce = job.c.clone(memory_format=torch.preserve_format)
                job.close()
                job.c.copy_(c_init)
                ck2 = job.verify()
                out["gemm_nccl_transport_bit_ok"] = ck2["bit_equal"] and api.dist_transport() == "NCCL all-gather" and bool(torch.equal(ce, job.c))
                job.one_sided = api.dist_register(job.a_loc, job.b_loc)
        out[f"gemm_{tag}_bit_ok"] = ck["bit_equal"]
        out[f"gemm_{tag}_resid_ok"] = ck["resid"] < 1e-14
        out[f"gemm_{tag}_same_kernel_ok"] = ck["kernel"] == ck["kernel_replay"]
        out[f"gemm_{tag}_kernel"] = ck["kernel"]
        err = full_product_check(job)
        out[f"gemm_{tag}_vs_full_single_gpu_product_ok"] = err < 1e-12
        out[f"gemm_{tag}_vs_full_err"] = err
        if tag == "4steps":
            # three products back to back (static A/B: the next product's first gather overlaps the current one)
            c0 = job.c.clone(memory_format=torch.preserve_format)
            for _ in range(3):
                job.step()
            torch.cuda.synchronize()
            got = job.c.clone(memory_format=torch.preserve_format)
            job.c.copy_(c0)
            for _ in range(3):
                job.step(); torch.cuda.synchronize()
            out["gemm_back_to_back_ok"] = bool(torch.equal(got, job.c))
            # A shard written on the stream right before the call: without the static flag the gather must wait for it
            job.c.copy_(c0)
            a_keep = job.a_loc.clone()
            job.a_loc.fill_(float("nan"))
            torch.cuda.synchronize()
            big = torch.empty(64 << 20, dtype=torch.float64, device=dev)
            for _ in range(4):
                big.normal_()                                  # keep the stream busy ahead of the copy
            job.a_loc.copy_(a_keep)
            job.step(flags=0); torch.cuda.synchronize()
            one = job.c.clone(memory_format=torch.preserve_format)
            job.c.copy_(c0); job.step(); torch.cuda.synchronize()
            out["gemm_inputs_written_on_stream_ok"] = bool(torch.equal(one, job.c)) and not bool(torch.isnan(one).any())
            del big
        job.close()
        del job
    # z through the same entry
    p = api.dist_plan(world, rank, pr * 512, pc * 384, 2 * 128 * L, 128)
    m_loc, n_loc = p.m1 - p.m0, p.n1 - p.n0
    g = torch.Generator(device=dev); g.manual_seed(77 + rank)
    cplx = lambda *s: torch.complex(torch.rand(*s, dtype=torch.float64, device=dev, generator=g) - 0.5,   # noqa: E731
                                    torch.rand(*s, dtype=torch.float64, device=dev, generator=g) - 0.5)
    za, zb, zc = cplx(p.na, 128, m_loc), cplx(p.nb, n_loc, 128), cplx(n_loc, m_loc)
    zc0 = zc.clone()
    api.dist_gemm(torch.complex128, pr * 512, pc * 384, 2 * 128 * L, 128, 2.0 + 0.5j, za, zb, 1.2 - 0.25j, zc.t(), 1, m_loc)
    torch.cuda.synchronize()
    row_pg = col_pg = None
    for i in range(pr):
        gq = dist.new_group([i * pc + jj for jj in range(pc)])
        if i == p.i: row_pg = gq
    for j in range(pc):
        gq = dist.new_group([ii * pc + j for ii in range(pr)])
        if j == p.j: col_pg = gq
    za_all = torch.empty((pc,) + tuple(za.shape), dtype=za.dtype, device=dev); zb_all = torch.empty((pr,) + tuple(zb.shape), dtype=zb.dtype, device=dev)
    dist.all_gather_into_tensor(torch.view_as_real(za_all).view(-1), torch.view_as_real(za).reshape(-1), group=row_pg)
    dist.all_gather_into_tensor(torch.view_as_real(zb_all).view(-1), torch.view_as_real(zb).reshape(-1), group=col_pg)
    a_rows = torch.cat([za_all[t % pc][t // pc] for t in range(p.T)], dim=0)
    b_cols = torch.cat([zb_all[t % pr][t // pr] for t in range(p.T)], dim=1)
    want = (1.2 - 0.25j) * zc0.t() + (2.0 + 0.5j) * (a_rows.t() @ b_cols.t())
    out["zgemm_err"] = float((zc.t() - want).abs().max())
    out["zgemm_ok"] = out["zgemm_err"] < 1e-12

    # ---------------------------------------------------------------- skinny, 1-D
    for dt, tagd in ((torch.float64, "d"), (torch.float32, "s")):
        job = bdist.DistSkinnyGemm(2048 + 64, world * 1536 + 200, 64, world, rank, dev, alpha=2.0, beta=1.2, root=0, dtype=dt)
        ck = job.verify()
        out[f"skinny_cols_{tagd}_bit_ok"] = ck["bit_equal"]
        out[f"skinny_cols_{tagd}_resid_ok"] = ck["resid"] < (1e-14 if dt == torch.float64 else 1e-5)
        del job
    # rows split: C's and A's rows local, B broadcast from the LAST rank
    m, n, k = world * 1024 + 77, 1920, 64
    i0, i1 = partition.thread_range_sub(rank, world, m, 128)
    g.manual_seed(5)
    b_full = (torch.rand(n, k, dtype=torch.float64, device=dev, generator=g) - 0.5)          # column-major k x n (same on every rank)
    g.manual_seed(6 + rank)
    a_l = torch.rand(k, max(i1 - i0, 1), dtype=torch.float64, device=dev, generator=g)[:, : i1 - i0].contiguous() - 0.5   # column-major m_loc x k
    c_l = torch.rand(n, max(i1 - i0, 1), dtype=torch.float64, device=dev, generator=g)[:, : i1 - i0].contiguous() - 0.5   # column-major m_loc x n
    c_ref = c_l.clone()
    b_buf = b_full.clone() if rank == world - 1 else torch.full_like(b_full, float("nan"))
    api.dist_gemm_1d(torch.float64, api.DIST_ROWS, world - 1, m, n, k, 2.0, a_l.t(), 1, i1 - i0, b_buf.t(), 1, k, 1.2, c_l.t(), 1, i1 - i0)
    torch.cuda.synchronize()
    if i1 > i0:
        api.bli_dgemm(0, 0, i1 - i0, n, k, 2.0, a_l.t(), 1, i1 - i0, b_full.t(), 1, k, 1.2, c_ref.t(), 1, i1 - i0)
        torch.cuda.synchronize()
    out["skinny_rows_bit_ok"] = bool(torch.equal(b_buf, b_full)) and bool(torch.equal(c_l, c_ref))

    # ---------------------------------------------------------------- trsm
    m, n = 1536 + 64, world * 640 + 130
    job = bdist.DistTrsm(m, n, world, rank, dev, alpha=2.0)
    a_keep = job.a.clone(memory_format=torch.preserve_format)
    if rank != 0:
        job.a.fill_(float("nan"))
    job.step(root=0); torch.cuda.synchronize()
    got = job.b.clone()
    ok_a = bool(torch.equal(job.a, a_keep))
    ref = job.b0.clone()
    if job.n_loc:
        api.bli_dtrsm(0, 0xC0, 0, 0, m, job.n_loc, 2.0, a_keep, 1, m, ref.t(), 1, m)
        torch.cuda.synchronize()
        tv = (torch.rand(job.n_loc, dtype=torch.float64, device=dev, generator=g) - 0.5) / n
        x, rhs = got.t() @ tv, 2.0 * (job.b0.t() @ tv)
        w = torch.linalg.solve_triangular(a_keep, rhs.unsqueeze(1), upper=False).squeeze(1)
        out["trsm_resid"] = float(torch.linalg.vector_norm(x - w) / max(1.0, float(torch.linalg.vector_norm(w))))
    else:
        out["trsm_resid"] = 0.0
    out["trsm_bit_ok"] = ok_a and bool(torch.equal(got, ref))
    out["trsm_resid_ok"] = out["trsm_resid"] < 1e-13
    out["launches"] = api.launch_count()
    dist.barrier()
    bdist.native_finalize()
finally:
    print(json.dumps(out), flush=True)
    dist.destroy_process_group()
