"""Variable store: the stand-in for TensorFlow's variable scopes on this path.

The reference creates variables lazily by name inside ``tf.variable_scope(name)``
(dyn_mult_view/mv3d/utils/tf_utils.py:60-65, 71-80, 88-95): ``name/Matrix``, ``name/b``
(FC), ``name/w``, ``name/b`` (conv), ``name/w`` (deconv).  The same names are kept here so a
checkpoint maps 1:1.

Layout in HBM: after the first forward has created every variable, ``finalize()`` moves the
fp32 masters, their gradients, Adam's m and v, and the bf16 compute copies into five flat
buffers in creation order.  Gradients are written straight into the flat gradient buffer by
the wgrad kernels (no autograd accumulation pass); data-parallel buckets are contiguous
slices of it, and Adam is one multi-tensor launch over the flat views.
"""
import contextlib
import math

import torch

_ALIGN = 64  # elements; keeps every view 128/256-byte aligned for vector and TMA access
BIG_VARIABLE = 1 << 20   # elements; variables this large (the FC matrices) begin a new exchange / Adam chunk group


class Variable:
    __slots__ = ("name", "shape", "master", "grad", "m", "v", "half", "trainable", "offset", "numel", "store", "used", "fused_adam", "rows")

    def __init__(self, name, tensor, trainable=True):
        self.name = name
        self.shape = tuple(tensor.shape)
        self.master = tensor
        self.numel = tensor.numel()
        self.grad = self.m = self.v = self.half = None
        self.trainable = trainable
        self.offset = -1
        self.used = False
        self.fused_adam = False     # optimizer.py: the weight gradient of this FC matrix is consumed by Adam inside one kernel
        self.rows = 0               # FC matrices: rows (samples) of the layer's input when it last ran forward


class VariableStore:
    def __init__(self, device, seed=0):
        self.device = torch.device(device)
        self.vars = {}          # name -> Variable, in creation order
        self._scope = []
        self.finalized = False
        self.gen = torch.Generator(device="cpu")
        self.gen.manual_seed(seed)
        self.grad_ready_hook = None     # called with a Variable when its gradient has been written
        self.flat = {}
        self.record = None              # dict: layer name -> activation, filled by tf_utils ops when set (tests)
        self.adam_live = None           # the optimizer, between its begin_step() and the end of backward (fused FC update)
        self.pre_use_hook = None        # called with a Variable before a layer reads it (deferred updates, data_parallel.py)
        self.prepack = None             # functional.Prepack: per-step weight packing on a side stream
        # dummy differentiable leaf: keeps the autograd tape alive for layers whose only
        # differentiable inputs are parameters (gradients of parameters bypass autograd)
        self.anchor = None
        self._pass_uses = set()
        self.new_anchor()

    def new_anchor(self):
        """A fresh leaf per forward pass.  Autograd pins a leaf's AccumulateGrad node to the
        stream that was current when the node was created and joins that stream at the end
        of backward; a leaf that outlives a step would tie a CUDA-graph capture to the
        (uncaptured) warm-up stream -- cudaErrorStreamCaptureIsolation."""
        self.anchor = torch.empty(1, device=self.device, requires_grad=True)
        self._pass_uses = set()
        if self.device.type == "cuda":
            from . import functional          # a backward pass that died mid-way never ran its side-stream join callback
            functional.reset_side_stream_state()
        return self.anchor

    # -- scopes -------------------------------------------------------------------------
    @contextlib.contextmanager
    def scope(self, name):
        self._scope.append(name)
        try:
            yield
        finally:
            self._scope.pop()

    def full_name(self, name):
        return "/".join(self._scope + [name])

    # -- creation (reference initialisers, tf_utils.py:54-98) -----------------------------
    def get(self, name, shape, init, stddev=0.0):
        full = self.full_name(name)
        v = self.vars.get(full)
        if v is not None:
            if tuple(shape) != v.shape:
                raise ValueError("variable %s exists with shape %s, requested %s" % (full, v.shape, tuple(shape)))
            # the weight-gradient kernels OVERWRITE the variable's slot of the flat gradient buffer and report it to
            # the data-parallel hook once: a second use of a name within one forward pass (TF-style reuse) would keep
            # only the last use's gradient -- refuse it instead of training on a wrong gradient
            if full in self._pass_uses:
                raise RuntimeError("variable %s is used twice in one forward pass; weight sharing across calls is not "
                                   "supported (fold the shared applications into the batch dimension)" % full)
            self._pass_uses.add(full)
            v.used = True
            return v
        if self.finalized:
            raise RuntimeError("variable store is finalized; cannot create %s" % full)
        shape = tuple(int(s) for s in shape)
        if self.device.type == "meta":         # shape-only build (CPU host-logic tests)
            v = Variable(full, torch.empty(shape, dtype=torch.float32, device="meta"))
            v.used, v.store = True, self
            v.grad = torch.empty(shape, dtype=torch.float32, device="meta")
            v.half = torch.empty(shape, dtype=torch.bfloat16, device="meta")
            self.vars[full] = v
            self._pass_uses.add(full)
            return v
        if init == "zeros":
            t = torch.zeros(shape, dtype=torch.float32)
        elif init == "normal":                 # tf.random_normal_initializer(stddev)
            t = torch.randn(shape, generator=self.gen, dtype=torch.float32) * stddev
        elif init == "truncated_normal":       # tf.truncated_normal_initializer: re-draw beyond 2 sigma
            t = torch.randn(shape, generator=self.gen, dtype=torch.float32)
            bad = t.abs() > 2
            while bool(bad.any()):
                t[bad] = torch.randn(int(bad.sum()), generator=self.gen, dtype=torch.float32)
                bad = t.abs() > 2
            t = t * stddev
        else:
            raise ValueError(init)
        v = Variable(full, t.to(self.device))
        v.used = True
        v.store = self
        # stand-alone buffers until finalize() re-homes them into the flat storage
        v.grad = torch.zeros(shape, dtype=torch.float32, device=self.device)
        v.half = torch.empty(shape, dtype=torch.bfloat16, device=self.device)
        self._cast(v.master, v.half, v.numel)
        self.vars[full] = v
        self._pass_uses.add(full)
        return v

    def _cast(self, src, dst, n):
        from . import _lib
        if self.device.type != "cuda":
            raise _lib.DmvError("variables live on a CUDA device; there is no CPU path")
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.call("dmv_cast_f32_to_bf16", src.data_ptr(), dst.data_ptr(), n, st)

    # -- flat storage -----------------------------------------------------------------------
    def finalize(self, trainable=None):
        """Flatten.  ``trainable`` is an optional predicate name -> bool; variables that never
        receive a gradient (the dead FCs of highdim_angle.py:8-9) must be excluded."""
        if self.finalized:
            return
        # Weights first (creation order), biases last: the fp32 bias masters are read directly by the conv / linear
        # epilogues, so under sharded data parallelism (data_parallel.py) they must stay replicated on every rank;
        # keeping them in one contiguous tail region makes that a single small allreduce.
        total = 0
        for v in self.vars.values():
            if trainable is not None:
                v.trainable = bool(trainable(v.name))
        for v in self.vars.values():
            if not v.name.endswith("/b"):
                if v.numel >= BIG_VARIABLE:          # big matrices start on a chunk-plan boundary (data_parallel.plan_chunks)
                    total = -(-total // 16384) * 16384
                v.offset = total
                total += int(math.ceil(v.numel / _ALIGN)) * _ALIGN
        self.shard_end = total = -(-total // 16384) * 16384
        for v in self.vars.values():
            if v.name.endswith("/b"):
                v.offset = total
                total += int(math.ceil(v.numel / _ALIGN)) * _ALIGN
        dev = self.device
        self.total = total
        # the buffers are allocated to a multiple of 16384 elements so that data-parallel shards (world x 256
        # elements, data_parallel.py) tile them exactly; the tail stays zero
        self.alloc = alloc = -(-total // 16384) * 16384
        if dev.type == "meta":
            self.finalized = True
            return
        self.flat = {
            "master": torch.zeros(alloc, dtype=torch.float32, device=dev),
            "grad": torch.zeros(alloc, dtype=torch.float32, device=dev),
            "m": torch.zeros(alloc, dtype=torch.float32, device=dev),
            "v": torch.zeros(alloc, dtype=torch.float32, device=dev),
            "half": torch.zeros(alloc, dtype=torch.bfloat16, device=dev),
        }
        for v in self.vars.values():
            sl = slice(v.offset, v.offset + v.numel)
            self.flat["master"][sl].copy_(v.master.reshape(-1))
            v.master = self.flat["master"][sl].view(v.shape)
            v.grad = self.flat["grad"][sl].view(v.shape)
            v.m = self.flat["m"][sl].view(v.shape)
            v.v = self.flat["v"][sl].view(v.shape)
            v.half = self.flat["half"][sl].view(v.shape)
        self.total = total
        self.finalized = True
        self.refresh_half()

    def refresh_half(self):
        """bf16 compute copies of every master (Adam keeps them current afterwards)."""
        self._cast(self.flat["master"], self.flat["half"], self.total)

    def trainable_vars(self):
        return [v for v in self.vars.values() if v.trainable]

    def notify_grad(self, var):
        if self.grad_ready_hook is not None:
            self.grad_ready_hook(var)

    # -- checkpoint surface (names as in a TF-1.3 checkpoint) --------------------------------
    def state_dict(self):
        return {k: v.master.detach().cpu().clone() for k, v in self.vars.items()}

    def load_state_dict(self, sd):
        for k, t in sd.items():
            v = self.vars[k]
            v.master.copy_(torch.as_tensor(t).to(self.device).reshape(v.shape))
            if not self.finalized:
                self._cast(v.master, v.half, v.numel)
        if self.finalized:
            self.refresh_half()

    def num_params(self):
        return sum(v.numel for v in self.vars.values())


_current = []


def current_store():
    if not _current:
        raise RuntimeError("no active VariableStore; wrap the graph in `with use_store(store):`")
    return _current[-1]


@contextlib.contextmanager
def use_store(store):
    _current.append(store)
    try:
        yield store
    finally:
        _current.pop()
