"""TF-flavoured Adam over the flat variable storage (replaces
tf.train.AdamOptimizer(lr).minimize, appearance_flow_model.py:77)."""
import ctypes as C
import os

import torch

from . import _lib

FUSE_MIN = 1 << 20      # elements: FC matrices this large take the weight-gradient + Adam kernel (dmv_linear_wgrad_adam)


def fusable(v):
    """An FC matrix whose weight gradient dmv_linear_wgrad_adam can consume in place (include/dmv3d.h), fed by at most 64
    samples: the streaming form of the kernel holds one 64-sample operand tile (with the 4 source frames of config 5 folded
    into the batch, 256 rows, the generic form lost 3 % against the separate calls)."""
    return (v.trainable and v.name.endswith("/Matrix") and len(v.shape) == 2 and v.numel >= FUSE_MIN and v.shape[0] % 8 == 0
            and v.shape[1] % 8 == 0 and v.numel % 256 == 0 and v.offset % 256 == 0 and 0 < v.rows <= 64)


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
        # Single-process training, optional (DMV_FUSE_FC_ADAM): big FC matrices are updated by the kernel that forms their
        # weight gradient, during backward (functional._Linear.backward) -- the gradient is never written.  The model's
        # train_step brackets backward with begin_step() / step(); a bare loss.backward() still writes plain gradients.
        # data_parallel.attach() clears the marks when the gradients must be exchanged first (world > 1).
        # MEASURED (profiles/r02_fc_adam_512.txt, r02_fuse_defer_sweep.txt): the fused kernel moves 26 instead of 34
        # B/parameter and runs at 0.88 of the HBM peak: 640 us for the four matrices against 827 us for the two calls.
        # Inside the step the best split is fc1 / a3 / a4 fused and a5 updated by the plain Adam chunks at the start of the
        # next step (deferred, data_parallel.py): 2.58 ms, against 2.63 with all four fused and 2.61 with fc1 alone.
        self._ticked = False
        # "auto": every eligible matrix but the last-created one (a5: its gradient is the first to arrive, and its plain Adam
        # chunks are deferred into the next forward pass, data_parallel.py) | "0" | "1" (every eligible matrix) | names
        fuse = os.environ.get("DMV_FUSE_FC_ADAM", "auto")
        if fuse != "0" and store.device.type == "cuda":
            cand = [v for v in self.vars if fusable(v)]
            for v in cand:
                v.fused_adam = (fuse == "1") or (fuse == "auto" and (v is not cand[-1] or len(cand) == 1)) or (v.name in fuse.split(","))
        # contiguous runs of trainable variables collapse into single tensors (the flat
        # buffers are contiguous), which keeps the launch count at one in the common case
        self._coalesce()

    def _coalesce(self):
        self._tbl = self._table([v for v in self.vars if not v.fused_adam])
        # the fused FC matrices, for a step() that was not bracketed by begin_step(): their plain gradients were written
        self._fused_tbl = self._table([v for v in self.vars if v.fused_adam])

    def _table(self, vs):
        runs = []
        for v in sorted(vs, key=lambda u: u.offset):
            pad_end = v.offset + -(-v.numel // 64) * 64
            if runs and runs[-1][1] == v.offset:
                runs[-1][1] = pad_end
            else:
                runs.append([v.offset, pad_end])
        f = self.store.flat
        n = len(runs)
        vp = C.c_void_p * n
        es = 4
        return (vp(*[f["master"].data_ptr() + a * es for a, _ in runs]), vp(*[f["grad"].data_ptr() + a * es for a, _ in runs]),
                vp(*[f["m"].data_ptr() + a * es for a, _ in runs]), vp(*[f["v"].data_ptr() + a * es for a, _ in runs]),
                vp(*[f["half"].data_ptr() + a * 2 for a, _ in runs]), (C.c_longlong * n)(*[b - a for a, b in runs]), n)

    def disable_fusion(self):
        for v in self.vars:
            v.fused_adam = False
        self._coalesce()

    def fused_vars(self):
        return [v for v in self.vars if v.fused_adam]

    def tick(self):
        from . import functional as F
        st = torch.cuda.current_stream(self.store.device).cuda_stream
        F._tag[0] = "adam"
        F.call("dmv_adam_tick", self.state.data_ptr(), self.lr, self.beta1, self.beta2, st)

    def begin_step(self):
        """Before backward: advance the step scalars and arm the in-backward update of the fused FC matrices."""
        if self.fused_vars():
            self.tick()
            self._ticked = True
            self.store.adam_live = self

    def end_backward(self):
        self.store.adam_live = None

    def step(self):
        from . import functional as F
        st = torch.cuda.current_stream(self.store.device).cuda_stream
        self.store.adam_live = None
        tables = [self._tbl]
        if not self._ticked:          # plain backward + step(): every gradient, the FC matrices' included, is in memory
            self.tick()
            tables.append(self._fused_tbl)
        self._ticked = False
        F._tag[0] = "adam"
        for p, g, m, v, h, n, count in tables:
            if count:
                F.call("dmv_adam_multi", p, g, m, v, h, n, count, self.state.data_ptr(), self.beta1, self.beta2, self.eps, self.grad_scale, st)

    @property
    def t(self):
        return int(self.state[3].item())
