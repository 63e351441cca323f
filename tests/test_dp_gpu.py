"""GPU tests of the data-parallel exchange kernel (csrc/exchange.cu) and, on a box with >= 2 GPUs, of the data-parallel
modes end to end (tools/check_dp.py under torchrun)."""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_exchange_kernel_emulated_ranks_equal_allreduce_then_adam(world):
    """``world`` ranks emulated on ONE device (every rank's buffers live on cuda:0, every rank's kernel on its own
    stream, the kernels wait on each other through the signal words exactly as across GPUs): after two steps every
    rank's bf16 buffer, and the owners' theta/m/v slices, are BIT-equal to summing the gradients in rank order and
    running dmv_adam_multi -- the allreduce mode's arithmetic."""
    from dynamic_multiview_3d_b200 import _lib
    dev = torch.device("cuda:0")
    L = _lib.load()
    n_slice, chunks = 4096 + 512, 3
    chunk = n_slice * world
    total = chunk * chunks
    tail = 256
    alloc = total + tail
    gen = torch.Generator(device=dev).manual_seed(world)
    words = L.dmv_dp_signal_words(chunks + 1)
    init = torch.randn(alloc, device=dev, generator=gen)
    ranks = []
    for r in range(world):
        ranks.append({"grad": torch.empty(alloc, device=dev), "half": torch.zeros(alloc, dtype=torch.bfloat16, device=dev),
                      "master": init.clone(), "m": torch.zeros(alloc, device=dev), "v": torch.zeros(alloc, device=dev),
                      "sig": torch.zeros(words, dtype=torch.int32, device=dev), "local": torch.zeros(2 * (chunks + 1), dtype=torch.int32, device=dev),
                      "stream": torch.cuda.Stream(device=dev)})
    vp = C.c_void_p * world
    gp, hp, sp = vp(*[x["grad"].data_ptr() for x in ranks]), vp(*[x["half"].data_ptr() for x in ranks]), vp(*[x["sig"].data_ptr() for x in ranks])
    # reference: replicated state, gradients summed in rank order, one dmv_adam_multi per step
    ref = {"master": init.clone(), "m": torch.zeros(alloc, device=dev), "v": torch.zeros(alloc, device=dev),
           "half": torch.zeros(alloc, dtype=torch.bfloat16, device=dev)}
    state = torch.tensor([1.0, 1.0, 0.0, 0.0], device=dev)
    one = C.c_void_p * 1
    for step in range(2):
        grads = [torch.randn(alloc, device=dev, generator=gen) * (1 + r) for r in range(world)]
        gsum = grads[0].clone()
        for r in range(1, world):
            gsum = gsum + grads[r]
        for r in range(world):
            ranks[r]["grad"].copy_(grads[r])
        st = torch.cuda.current_stream().cuda_stream
        _lib.call("dmv_adam_tick", state.data_ptr(), 1e-3, 0.9, 0.999, st)
        _lib.call("dmv_adam_multi", one(ref["master"].data_ptr()), one(gsum.data_ptr()), one(ref["m"].data_ptr()), one(ref["v"].data_ptr()),
                  one(ref["half"].data_ptr()), (C.c_longlong * 1)(alloc), 1, state.data_ptr(), 0.9, 0.999, 1e-8, 1.0, st)
        torch.cuda.synchronize()
        for c in list(range(chunks)) + [-1]:
            for r in range(world):                                 # launch order differs per rank pair on purpose: r ascending
                x = ranks[r]
                start, n, slot, rep = (c * chunk, n_slice, c, 0) if c >= 0 else (total, tail, chunks, 1)
                _lib.call("dmv_dp_exchange_chunk", gp, hp, sp, None, None, x["master"].data_ptr(), x["m"].data_ptr(), x["v"].data_ptr(),
                          x["local"].data_ptr(), start, n, r, world, slot, rep, state.data_ptr(), 0.9, 0.999, 1e-8, 1.0, 4, x["stream"].cuda_stream)
        torch.cuda.synchronize()
        for r in range(world):
            x = ranks[r]
            assert torch.equal(x["half"].view(torch.int16), ref["half"].view(torch.int16)), "bf16 copy of rank %d differs (step %d)" % (r, step)
            for c in range(chunks):
                a = c * chunk + r * n_slice
                for k in ("master", "m", "v"):
                    assert torch.equal(x[k][a:a + n_slice], ref[k][a:a + n_slice]), (k, r, c, step)
            for k in ("master", "m", "v"):                         # the replicated tail is updated by every rank
                assert torch.equal(x[k][total:], ref[k][total:]), (k, r, "tail", step)
    assert int(ranks[0]["local"][0]) == 2 and int(ranks[0]["local"][1]) == 0          # epoch advanced, ticket reset


def test_exchange_kernel_rejects_bad_arguments():
    from dynamic_multiview_3d_b200 import _lib
    L = _lib.load()
    t = torch.zeros(64, device="cuda")
    vp = C.c_void_p * 3
    p = vp(t.data_ptr(), t.data_ptr(), t.data_ptr())
    assert L.dmv_dp_exchange_chunk(p, p, p, None, None, t.data_ptr(), t.data_ptr(), t.data_ptr(), t.data_ptr(), 0, 8, 0, 3, 0, 0, t.data_ptr(),
                                   0.9, 0.999, 1e-8, 1.0, 0, None) == -3          # world 3: unsupported shape
    assert L.dmv_dp_exchange_chunk(p, p, p, None, None, t.data_ptr(), t.data_ptr(), t.data_ptr(), t.data_ptr(), 0, 12, 0, 2, 0, 0, t.data_ptr(),
                                   0.9, 0.999, 1e-8, 1.0, 0, None) == -2          # n_slice % 8
    assert L.dmv_dp_signal_words(5) == 80


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs on the box")
def test_data_parallel_modes_agree_across_gpus():
    """allreduce == sharded (NCCL) == fused (one kernel per chunk over peer memory), replicas identical on every rank,
    fused mode bit-reproducible (tools/check_dp.py under torchrun)."""
    n = min(torch.cuda.device_count(), 8)
    n = 1 << (n.bit_length() - 1)
    port = str(29700 + os.getpid() % 200)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                          "--master-port", port, os.path.join(ROOT, "tools", "check_dp.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "check_dp ok" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
