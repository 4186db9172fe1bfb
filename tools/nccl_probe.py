"""NCCL subgroup all-gather bandwidth probe (dev tool). torchrun --nproc-per-node N tools/nccl_probe.py"""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
pc = 2 if world == 4 else (4 if world == 8 else 1); pr = world // pc
i, j = divmod(rank, pc)
row = col = None
for ii in range(pr):
    g = dist.new_group([ii * pc + jj for jj in range(pc)])
    if ii == i: row = g
for jj in range(pc):
    g = dist.new_group([ii * pc + jj for ii in range(pr)])
    if jj == j: col = g
for name, grp, n in (("world", None, world), ("row", row, pc), ("col", col, pr)):
    if n == 1: continue
    src = torch.ones(128 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)        # 128 MiB per rank
    dst = torch.empty(n * src.numel(), dtype=torch.float64, device=dev)
    for _ in range(3): dist.all_gather_into_tensor(dst, src, group=grp)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): dist.all_gather_into_tensor(dst, src, group=grp)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    if rank == 0: print(f"{name}: n={n} all_gather 128MiB/rank {dt*1e3:.2f} ms  recv {(n-1)*128/1024/dt:.1f} GiB/s", flush=True)
# peer copy bandwidth (cudaMemcpyPeer through torch)
if world >= 2 and rank == 0:
    a = torch.ones(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda:0"); b = torch.empty_like(a, device="cuda:1")
    b.copy_(a); torch.cuda.synchronize()
    t0 = time.perf_counter(); b.copy_(a); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"peer copy 256MiB cuda:0->cuda:1 {dt*1e3:.2f} ms = {0.25/dt:.1f} GiB/s; can_access_peer={torch.cuda.can_device_access_peer(0,1)}", flush=True)
dist.destroy_process_group()
