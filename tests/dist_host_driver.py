"""Multi-GPU gemm with the shards in pinned host memory (DistGemm.step_host) against the device-resident product
(DistGemm.step) on the same shards.  Splitting a k step into column blocks does not change any element's summation
order, so the two agree bit for bit whenever the engine picks the same kernel for the narrower blocks (measured: N=2,
both cases); when it picks another tile shape (small problems) the epilogue may round beta*C + alpha*AB differently by
one ulp -- the bar is therefore 1e-14 absolute on O(1) data, and bit-exactness is reported.  Run under torchrun (any
world size) or stand-alone (world size 1); prints one JSON line per rank."""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from blis_b200 import dist as bdist            # noqa: E402

rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
out = {"rank": rank, "world": world}
try:
    for tag, (n, k, kb) in {"4steps": (1536, 4096, 512), "1step": (1024, 0, 512)}.items():
        pr, pc = bdist.partition.thread_partition_2x2(world, n, n)
        import math
        L = math.lcm(pr, pc)
        k = max(k, kb * L)                      # "1step": exactly one k step (first == last)
        job = bdist.DistGemm(pr * n, pc * n, k, world, rank, dev, alpha=2.0, beta=1.2, kb=kb)
        hosts = job.host_shards()
        c0 = job.c.clone(memory_format=torch.preserve_format)
        job.step(); torch.cuda.synchronize()
        want = job.c.t().clone()
        # poison everything on the device: step_host must bring all of it from the host images
        job.a_loc.fill_(float("nan")); job.b_loc.fill_(float("nan")); job.c.fill_(float("nan"))
        for rep in range(2):
            hosts[2].copy_(c0.t())
            job.step_host(hosts, nblk=3); torch.cuda.synchronize()
            out[f"{tag}_rep{rep}_bit_exact"] = bool(torch.equal(hosts[2], want.cpu()))
            out[f"{tag}_rep{rep}_maxdiff"] = float((hosts[2] - want.cpu()).abs().max())
        out[f"{tag}_steps"] = job.plan.steps
    dist.barrier()
finally:
    print(json.dumps(out), flush=True)
    dist.destroy_process_group()
