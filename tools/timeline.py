"""Kernel timeline of the captured train step (CUPTI activity records through torch.profiler; no nsys in the image).
Replays the CUDA graph of bench config 2 a few times under the profiler and writes, for ONE replay, every kernel with
its stream, start (us from the first kernel of the step) and duration -- plus per-stream busy time and the gaps of the
main stream.  Timestamps under CUPTI tracing carry a little overhead; the per-kernel durations and the overlap structure
are what this is for (which stream is the critical path, what the tail of the step is)."""
import json
import os
import sys
import tempfile

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dynamic_multiview_3d_b200 as pkg  # noqa: E402
from dynamic_multiview_3d_b200 import data_parallel  # noqa: E402
from dynamic_multiview_3d_b200.train import GraphedTrainStep, synthetic_batch  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.txt"
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:                      # under torchrun: the data-parallel step (rank 0 writes its timeline)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    conf = {"batch_size": 64, "learning_rate": 1e-4, "image_size": 224, "viewpoint_dim": 19, "loss": "l2", "seed": 0}
    model = pkg.AppearanceFlowModel(conf)
    if world > 1 or os.environ.get("DMV_OVERLAP_ADAM", "1") == "1":
        data_parallel.attach(model, bucket_mb=float(os.environ.get("DMV_DP_CHUNK_MB", "128" if world > 1 else "32")))
    b = synthetic_batch(model, seed=1234, rank=rank)
    devb = {k: torch.from_numpy(v).to(dev) for k, v in b.items()}
    step = GraphedTrainStep(model, warmup=2)
    step(devb)
    for _ in range(5):
        step.replay()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step.replay()
        torch.cuda.synchronize()
    if rank != 0:
        torch.cuda.synchronize()
        os._exit(0)
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and e.get("ph") == "X"]
    ev.sort(key=lambda e: e["ts"])
    # split into replays at the sampler kernel of the fused warp + loss (one per step)
    marks = [i for i, e in enumerate(ev) if "adam_tick" in e["name"]]
    if len(marks) >= 3:
        # a step's first kernel follows the previous step's last; cut at the largest gap before each tick
        pass
    # simpler: cut by count
    n = len(ev) // 3
    one = ev[n:2 * n]
    t0 = one[0]["ts"]
    streams = sorted({e["args"].get("stream", -1) for e in one})
    with open(out, "w") as f:
        span = max(e["ts"] + e["dur"] for e in one) - t0
        f.write("# %d kernels in the replay, span %.1f us; streams %s\n" % (len(one), span, streams))
        for s in streams:
            es = [e for e in one if e["args"].get("stream", -1) == s]
            busy = sum(e["dur"] for e in es)
            f.write("# stream %s: %d kernels, busy %.1f us, first %.1f, last end %.1f\n"
                    % (s, len(es), busy, es[0]["ts"] - t0, max(e["ts"] + e["dur"] for e in es) - t0))
        f.write("# start_us dur_us stream grid kernel\n")
        for e in one:
            name = e["name"].replace("void ", "").replace("dmv::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
            name = name.split("(")[0]
            f.write("%9.1f %8.1f %3s %6s %s\n" % (e["ts"] - t0, e["dur"], e["args"].get("stream", -1), str(e["args"].get("grid", "")).replace(" ", ""), name[:70]))
    print(open(out).read()[:1500])
    if world > 1:
        sys.stdout.flush()
        os._exit(0)                    # a process group whose collectives live in a captured graph can block in destroy


if __name__ == "__main__":
    main()
