"""The reference path's train step on host cores -- TEST/BASELINE INFRASTRUCTURE ONLY.

TensorFlow 1.3 cannot run here, so the "reference CPU path" is the
oracle graph (oracle/graph.py) on its torch-CPU backend: fp32 forward + backward + the
TF-flavoured Adam of oracle/tf_ops.py, with every host thread torch can use.  bench.py times
it as ``cpu_baseline`` / ``--impl reference`` (kind "port").
"""
import time

import numpy as np
import torch

from . import graph as G


class CpuAppFlowStep:
    def __init__(self, H=224, V=19, kind="base", lr=1e-4, seed=0, threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.ops = G.TorchCpuOps()
        self.kind = kind
        shapes = G.appflow_param_shapes(H, V, kind)
        self.P = {k: torch.tensor(v, requires_grad=True) for k, v in G.init_params(shapes, seed).items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.P.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.P.items()}
        self.t, self.lr = 0, lr
        self.b1p = self.b2p = 1.0

    def step(self, image0, image1, disp, mode="l2"):
        out = G.appearance_flow_forward(self.ops, self.P, image0, disp, self.kind)
        loss = G.appearance_flow_loss(self.ops, out, image1, mode)
        grads = torch.autograd.grad(loss, [p for p in self.P.values()], allow_unused=True)
        self.t += 1
        self.b1p *= 0.9
        self.b2p *= 0.999
        lr_t = self.lr * (1 - self.b2p) ** 0.5 / (1 - self.b1p)
        with torch.no_grad():
            for (k, p), g in zip(self.P.items(), grads):
                if g is None:
                    continue
                m, v = self.m[k], self.v[k]
                m.add_((g - m) * 0.1)
                v.add_((g * g - v) * 0.001)
                p.sub_((m * lr_t) / (v.sqrt() + 1e-8))
        return float(loss)


def time_steps(batch, H=224, V=19, steps=3, warmup=1, seed=1234):
    """Returns (samples_per_s, seconds_per_step list, cores)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from dynamic_multiview_3d_b200.synthetic import make_batch
    b = make_batch(batch, H, "onehot19" if V == 19 else "disp2", seed=seed)
    st = CpuAppFlowStep(H, V)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        st.step(b["image0"], b["image1"], b["disp"])
        dt = time.perf_counter() - t0
        if i >= warmup:
            ts.append(dt)
    return batch / float(np.median(ts)), ts, torch.get_num_threads()
