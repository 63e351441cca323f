"""torchrun --nproc-per-node N tools/time_comm.py : raw NCCL time of one step's exchange (no compute running):
reduce-scatter of 555 MB fp32 in 32 MB chunks + all-gather of 278 MB bf16, and the plain allreduce."""
import os
import sys

import torch
import torch.distributed as dist

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
T = 138_756_096 // 16384 * 16384
g = torch.randn(T, device=dev)
h = torch.zeros(T, dtype=torch.bfloat16, device=dev)
chunk = int(os.environ.get("CHUNK_MB", 32)) * (1 << 20) // 4
chunks = [(s, min(T, s + chunk)) for s in range(0, T, chunk)]


def rs():
    for s, e in chunks:
        n = (e - s) // world
        dist.reduce_scatter_tensor(g[s + rank * n:s + (rank + 1) * n], g[s:e])


def ag():
    for s, e in chunks:
        n = (e - s) // world
        dist.all_gather_into_tensor(h[s:e], h[s + rank * n:s + (rank + 1) * n])


def ar():
    for s, e in chunks:
        dist.all_reduce(g[s:e])


for name, fn in (("reduce_scatter fp32", rs), ("all_gather bf16", ag), ("all_reduce fp32", ar)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print("world %d chunk %d MB  %-20s %.3f ms per step" % (world, chunk * 4 >> 20, name, e0.elapsed_time(e1) / 10), flush=True)
dist.barrier()
os._exit(0)
