"""Bisect which part of the train step invalidates CUDA-graph capture (run on the GPU box)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = ["fwd", "fwd_loss", "fwd_loss_bwd", "full", "full_threadlocal", "full_relaxed", "adam_only", "sampler_only", "conv_only"]


def run_case(case):
    import torch
    import dynamic_multiview_3d_b200 as pkg
    from dynamic_multiview_3d_b200 import functional as F
    from dynamic_multiview_3d_b200.synthetic import make_batch
    dev = torch.device("cuda:0")
    conf = {"batch_size": 2, "learning_rate": 1e-4, "image_size": 32, "viewpoint_dim": 19}
    m = pkg.AppearanceFlowModel(conf)
    b = make_batch(2, 32, "onehot19")
    i0, i1, d = (torch.from_numpy(b[k]).to(dev) for k in ("image0", "image1", "disp"))

    def body():
        if case == "fwd":
            with torch.no_grad():
                return m.forward(i0, d)["gen"]
        if case == "fwd_loss":
            with torch.no_grad():
                m.forward(i0, d)
                return m.build_loss(i1)
        if case == "fwd_loss_bwd":
            m.forward(i0, d)
            l = m.build_loss(i1)
            l.backward()
            return l.detach()
        if case == "adam_only":
            m.optimizer.step()
            return m.optimizer.state
        if case == "sampler_only":
            return F.flow_resampler(i0, torch.zeros(2, 32, 32, 2, device=dev))
        if case == "conv_only":
            with torch.no_grad():
                v = m.store.vars
                return F.conv2d(i0, v["e0/w"], v["e0/b"], 2, "lrelu")
        return m.train_step(i0, i1, d)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    mode = {"full_threadlocal": "thread_local", "full_relaxed": "relaxed"}.get(case, "global")
    with torch.cuda.graph(g, capture_error_mode=mode):
        out = body()
    g.replay()
    torch.cuda.synchronize()
    print("CASE", case, "OK", float(out.float().sum()))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for c in CASES:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), c], capture_output=True, text=True, timeout=300)
            tail = (r.stdout + r.stderr).strip().splitlines()[-3:]
            print("=== %s rc=%d :: %s" % (c, r.returncode, " | ".join(tail)), flush=True)
