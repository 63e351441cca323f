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


class CpuTrainStep:
    """fp32 forward + autograd backward + TF-flavoured Adam (oracle/tf_ops.py:adam rule; fp32 ``1 - beta`` as TF's
    ApplyAdam computes it) of one of the oracle graphs on the torch-CPU backend."""

    def __init__(self, shapes, loss_fn, lr=1e-4, seed=0, threads=None):
        if threads:
            torch.set_num_threads(threads)
        self.ops = G.TorchCpuOps()
        self.loss_fn = loss_fn                 # (ops, P, batch) -> scalar loss tensor
        self.P = {k: torch.tensor(v, requires_grad=True) for k, v in G.init_params(shapes, seed).items()}
        self.m = {k: torch.zeros_like(v) for k, v in self.P.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.P.items()}
        self.t, self.lr = 0, lr
        self.b1p = self.b2p = 1.0

    def load(self, params):
        """Replace the parameters (name -> array) and reset the optimizer state."""
        assert sorted(params) == sorted(self.P), "variable tables differ"
        self.P = {k: torch.tensor(np.asarray(params[k], np.float32), requires_grad=True) for k in self.P}
        self.m = {k: torch.zeros_like(v) for k, v in self.P.items()}
        self.v = {k: torch.zeros_like(v) for k, v in self.P.items()}
        self.t, self.b1p, self.b2p = 0, 1.0, 1.0

    def step_batch(self, batch):
        loss = self.loss_fn(self.ops, self.P, batch)
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
        return float(loss.detach())


class CpuAppFlowStep(CpuTrainStep):
    """The single-view appearance-flow step (BASELINE configs 1-3)."""

    def __init__(self, H=224, V=19, kind="base", lr=1e-4, seed=0, threads=None, mode="l2"):
        self.kind = kind

        def loss_fn(ops, P, b):
            out = G.appearance_flow_forward(ops, P, b["image0"], b["disp"], kind)
            return G.appearance_flow_loss(ops, out, b["image1"], mode)
        super().__init__(G.appflow_param_shapes(H, V, kind), loss_fn, lr, seed, threads)

    def step(self, image0, image1, disp):
        return self.step_batch({"image0": image0, "image1": image1, "disp": disp})


def make_config_step(config, H=224, V=19):
    """(CpuTrainStep, batch maker (B) -> dict of NumPy arrays, description) for a BASELINE config number (bench.py --config)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from dynamic_multiview_3d_b200 import synthetic as S
    if config in (1, 2):
        return CpuAppFlowStep(H, V, "base"), (lambda B: S.make_batch(B, H, "onehot19")), "single-view app-flow (appearance_flow_model.py:83-127), L2"
    if config == 3:
        return CpuAppFlowStep(H, V, "highdim"), (lambda B: S.make_batch(B, H, "onehot19")), "app-flow + offset head, high-dim viewpoint (highdim_angle.py:5-10), L2"
    if config == 4:
        conf = {"use_color": "", "use_depth": "", "depth_lr_factor": 0.1}

        def loss_fn(ops, P, b):
            out = G.colordepth_forward(ops, P, conf, b["image0"], b["depth0"], b["disp"])
            return G.colordepth_loss(ops, out, conf, b["image1"], b["depth1"], "l1")
        return (CpuTrainStep(G.colordepth_param_shapes(H, V, conf), loss_fn), (lambda B: S.make_batch(B, H, "onehot19", depth=True)),
                "cars_colordepth RGB+depth tanh heads (main_model.py:83-154), per-channel L1, weights (1, 0.1)")
    if config == 5:
        conf = {"use_color": "", "use_depth": 0.1}

        def loss_fn(ops, P, b):
            return G.multiview_loss(ops, G.multiview_forward(ops, P, conf, b), b["image1"])
        return (CpuTrainStep(G.multiview_param_shapes(H, V, conf), loss_fn),
                (lambda B: S.make_multiview_multiobject_batch(B, H, 4, viewpoint="onehot19")),
                "4 source frames on the multi-object trunk, 3-channel flow+confidence head, softmax fusion (SURVEY 8(f)-3), L2")
    raise ValueError(config)


def use_all_host_threads():
    """the CPU arm uses every host thread.  torchrun exports OMP_NUM_THREADS=1 to its workers,
    so the count is set explicitly instead of trusting the environment."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def time_steps(batch, H=224, V=19, steps=3, warmup=1, seed=1234, config=2):
    """Returns (samples_per_s, seconds_per_step list, cores)."""
    use_all_host_threads()
    st, make, _ = make_config_step(config, H, V)
    b = make(batch)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        st.step_batch(b)
        dt = time.perf_counter() - t0
        if i >= warmup:
            ts.append(dt)
    return batch / float(np.median(ts)), ts, torch.get_num_threads()
