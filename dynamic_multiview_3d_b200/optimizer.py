"""TF-flavoured Adam over the flat variable storage (replaces
tf.train.AdamOptimizer(lr).minimize, appearance_flow_model.py:77)."""
import ctypes as C

import torch

from . import _lib


class TFAdam:
    """m += (g-m)(1-b1); v += (g^2-v)(1-b2); theta -= m*lr_t/(sqrt(v)+eps),
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t) kept on the device so a captured graph replays as is.
    Also refreshes the bf16 compute copy of every parameter in the same pass."""

    def __init__(self, store, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
        if not store.finalized:
            raise RuntimeError("finalize the VariableStore before creating the optimizer")
        self.store = store
        self.lr, self.beta1, self.beta2, self.eps, self.grad_scale = float(lr), beta1, beta2, eps, float(grad_scale)
        self.state = torch.tensor([1.0, 1.0, 0.0, 0.0], dtype=torch.float32, device=store.device)
        self.vars = store.trainable_vars()
        n = len(self.vars)
        vp = C.c_void_p * n
        self._p = vp(*[v.master.data_ptr() for v in self.vars])
        self._g = vp(*[v.grad.data_ptr() for v in self.vars])
        self._m = vp(*[v.m.data_ptr() for v in self.vars])
        self._v = vp(*[v.v.data_ptr() for v in self.vars])
        self._h = vp(*[v.half.data_ptr() for v in self.vars])
        self._n = (C.c_longlong * n)(*[v.numel for v in self.vars])
        self._count = n
        # contiguous runs of trainable variables collapse into single tensors (the flat
        # buffers are contiguous), which keeps the launch count at one in the common case
        self._coalesce()

    def _coalesce(self):
        runs = []
        for v in sorted(self.vars, key=lambda u: u.offset):
            pad_end = v.offset + -(-v.numel // 64) * 64
            if runs and runs[-1][1] == v.offset:
                runs[-1][1] = pad_end
            else:
                runs.append([v.offset, pad_end])
        f = self.store.flat
        n = len(runs)
        vp = C.c_void_p * n
        es = 4
        self._p = vp(*[f["master"].data_ptr() + a * es for a, _ in runs])
        self._g = vp(*[f["grad"].data_ptr() + a * es for a, _ in runs])
        self._m = vp(*[f["m"].data_ptr() + a * es for a, _ in runs])
        self._v = vp(*[f["v"].data_ptr() + a * es for a, _ in runs])
        self._h = vp(*[f["half"].data_ptr() + a * 2 for a, _ in runs])
        self._n = (C.c_longlong * n)(*[b - a for a, b in runs])
        self._count = n

    def step(self):
        from . import functional as F
        st = torch.cuda.current_stream(self.store.device).cuda_stream
        F._tag[0] = "adam"
        F.call("dmv_adam_tick", self.state.data_ptr(), self.lr, self.beta1, self.beta2, st)
        F.call("dmv_adam_multi", self._p, self._g, self._m, self._v, self._h, self._n, self._count,
                  self.state.data_ptr(), self.beta1, self.beta2, self.eps, self.grad_scale, st)

    @property
    def t(self):
        return int(self.state[3].item())
