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
assert all((e - s) % (world * 256) == 0 for s, e in chunks)


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
# the same exchange as ONE hand-written kernel per chunk over peer memory (csrc/exchange.cu): reduce-scatter + Adam + all-gather
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import types  # noqa: E402

from dynamic_multiview_3d_b200 import _lib, data_parallel  # noqa: E402

for mc in ("0", "1"):
    os.environ["DMV_DP_MULTICAST"] = mc
    store = types.SimpleNamespace(device=dev, alloc=T, vars={}, flat={"grad": torch.randn(T, device=dev), "half": torch.zeros(T, dtype=torch.bfloat16, device=dev)})
    px = data_parallel.PeerExchange(store, len(chunks))
    if mc == "1" and not px.multicast:
        if rank == 0:
            print("no multicast mapping on this box: multimem variant skipped", flush=True)
        break
    master, m, v = torch.randn(T, device=dev), torch.zeros(T, device=dev), torch.zeros(T, device=dev)
    state = torch.tensor([0.9, 0.999, 1e-4, 1.0], device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def fused():
        for c, (s, e) in enumerate(chunks):
            _lib.call("dmv_dp_exchange_chunk", px.grad_peers, px.half_peers, px.sig_peers, px.grad_mc, px.half_mc, master.data_ptr(), m.data_ptr(),
                      v.data_ptr(), px.local.data_ptr(), s, (e - s) // world, rank, world, c, 0, state.data_ptr(), 0.9, 0.999, 1e-8, 1.0, px.ctas, st)

    for _ in range(3):
        fused()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fused()
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print("world %d chunk %d MB  %-20s %.3f ms per step (reduce-scatter + Adam on 1/%d + bf16 all-gather, ctas %s)" % (
            world, chunk * 4 >> 20, "fused kernel " + ("multimem" if px.multicast else "peer ld/st"), e0.elapsed_time(e1) / 10, world, px.ctas or "default"), flush=True)
dist.barrier()
os._exit(0)
