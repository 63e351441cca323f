"""torchrun --nproc-per-node N tools/check_dp.py : the data-parallel modes -- allreduce (replicated Adam), sharded (NCCL
reduce-scatter + owner Adam + all-gather) and fused (ONE hand-written kernel per chunk over peer memory: csrc/exchange.cu,
once with fixed-order peer loads and once with multimem in-switch reduction where the box offers a multicast mapping) --
must produce the same parameters, identical on every rank; the fused mode must be bit-reproducible run to run.
Also run by tests/test_dp_gpu.py when the box has >= 2 GPUs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import dynamic_multiview_3d_b200 as pkg  # noqa: E402
from dynamic_multiview_3d_b200 import data_parallel  # noqa: E402
from dynamic_multiview_3d_b200.synthetic import make_batch  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = {}
MODES = ("allreduce", "allreduce2", "sharded", "fused_p2p", "fused_p2p2", "fused_mc", "fused_mc2")
# Two ranks: a + b has one order, so every mode gives the same bits.  More ranks: NCCL's allreduce and reduce-scatter,
# the switch's multimem reduction and the kernel's rank-order sum each fix their own fp32 order; three Adam steps at
# lr 1e-3 turn that last-bit difference into ~1e-4 relative parameter difference (Adam's early updates are +-lr).
TOL = 1e-5 if world <= 2 else 2e-3
for mode in MODES:
    os.environ["DMV_DP_MULTICAST"] = "1" if mode.startswith("fused_mc") else "0"
    model = pkg.AppearanceFlowModel({"batch_size": 4, "learning_rate": 1e-3, "image_size": 64, "viewpoint_dim": 19, "seed": 0})
    red = data_parallel.attach(model, bucket_mb=4.0, mode=mode.split("_")[0].rstrip("2"))
    if mode.startswith("fused") and rank == 0:
        print("mode %s: reducer %s, multicast %s" % (mode, type(red).__name__, getattr(getattr(red, "px", None), "multicast", None)), flush=True)
    b = make_batch(4, 64, "onehot19", seed=7, rank=rank)
    args = [torch.from_numpy(b[k]).to(dev) for k in ("image0", "image1", "disp")]
    losses = [float(model.train_step(*args)) for _ in range(3)]
    sd = model.state_dict()
    res[mode + "_sd"] = {k: v.float() for k, v in sd.items() if not k.startswith("__")}
    res[mode] = (losses, torch.cat([sd[k].reshape(-1).float() for k in sorted(sd) if not k.startswith("__")]).to(dev))
    half = model.store.flat["half"][:model.store.total].float()
    ref = half.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(half, ref), "bf16 replicas differ across ranks in mode " + mode
la2, pa2 = res["allreduce2"]
print("rank %d run-to-run (allreduce twice) max rel param diff %.3g" % (rank, float((res["allreduce"][1] - pa2).abs().max() / pa2.abs().max())))
la, pa = res["allreduce"]
ls, ps = res["sharded"]
if rank == 0:
    rows = []
    for k, v in res["allreduce_sd"].items():
        w = res["sharded_sd"][k]
        rows.append((float((v - w).abs().max() / max(float(v.abs().max()), 1e-12)), k, tuple(v.shape)))
    for r in sorted(rows, reverse=True)[:12]:
        print("   diff %.3g %s %s" % r)
err = float((pa - ps).abs().max() / pa.abs().max())
print("rank %d losses allreduce %s sharded %s  max rel param diff %.3g" % (rank, la, ls, err))
assert err < TOL, err
for mode in ("fused_p2p", "fused_mc"):
    lf, pf = res[mode]
    errf = float((pa - pf).abs().max() / pa.abs().max())
    print("rank %d losses %s %s  max rel param diff vs allreduce %.3g" % (rank, mode, lf, errf))
    assert errf < TOL, (mode, errf)
assert torch.equal(res["fused_p2p"][1], res["fused_p2p2"][1]), "fused exchange is not bit-reproducible run to run"
print("rank %d fused (fixed-order peer loads) run-to-run: bit-identical" % rank)
print("rank %d fused (multimem in-switch reduction) run-to-run: %s" % (rank, "bit-identical" if torch.equal(res["fused_mc"][1], res["fused_mc2"][1]) else "DIFFERS"))
dist.barrier()
if rank == 0:
    print("check_dp ok")
os._exit(0)
