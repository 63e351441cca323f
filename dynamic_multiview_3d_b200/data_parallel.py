"""Batch-sharded data parallelism: one process per GPU, replicated parameters, ONE exchange
per step -- a sum-allreduce of the flat gradient buffer over NCCL (NVLink 5 / NVSwitch).
New work relative to the reference, which is single-GPU.

The flat gradient buffer (variables.py) is cut into contiguous buckets in REVERSE creation
order, i.e. the order backward produces gradients: decoder convs, a5, a4, a3, fc1, encoder.
When the last gradient of a bucket has been written, the bucket's allreduce is enqueued on a
side stream behind an event, so the four big FC buckets (97 % of the bytes) fly while the
FLOP-heavy encoder backward still runs.  The loss already divides by the GLOBAL pixel count,
so the summed gradient is the global-batch gradient and Adam needs no rescale.  Everything
is stream-ordered (events, no host sync), hence capturable in a CUDA graph.
"""
import torch
import torch.distributed as dist


def plan_buckets(var_table, bucket_elems):
    """var_table: [(name, offset, padded_numel)] in creation order -> buckets, last variable first.
    Each bucket is (start, end, [names]) with start/end offsets into the flat buffer."""
    buckets, names, start, end = [], [], None, None
    for name, off, n in reversed(var_table):
        if not names:
            end = off + n
        start = off
        names.append(name)
        if end - start >= bucket_elems:
            buckets.append((start, end, names))
            names = []
    if names:
        buckets.append((start, end, names))
    return buckets


class GradientAllReducer:
    def __init__(self, flat_grad, var_table, bucket_mb=32.0, group=None, trainable=None):
        """flat_grad: 1-D fp32 tensor; var_table as in plan_buckets; trainable: names that will
        actually report a gradient each step (others are not waited for)."""
        self.flat = flat_grad
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets = plan_buckets(var_table, int(bucket_mb * (1 << 20) / 4))
        self.var2bucket = {}
        self.expected = []
        for bi, (_, _, names) in enumerate(self.buckets):
            live = [n for n in names if trainable is None or n in trainable]
            self.expected.append(len(live))
            for n in live:
                self.var2bucket[n] = bi
        self.pending = list(self.expected)
        self.cuda = flat_grad.is_cuda
        self.comm_stream = torch.cuda.Stream(device=flat_grad.device) if self.cuda else None
        self.launched = []

    def on_grad_ready(self, var):
        bi = self.var2bucket.get(var.name if hasattr(var, "name") else var)
        if bi is None:
            return
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def _launch(self, bi):
        start, end, _ = self.buckets[bi]
        chunk = self.flat[start:end]
        self.launched.append(bi)
        if self.world == 1:
            return
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.flat.device))
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)

    def finish(self):
        """Join the side stream; buckets whose variables never reported (should not happen)
        are reduced now so the ranks stay consistent."""
        for bi, p in enumerate(self.pending):
            if p > 0 and self.expected[bi] > 0:
                self._launch(bi)
        if self.cuda and self.world > 1:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm_stream)
        self.pending = list(self.expected)
        self.launched = []


def attach(model, bucket_mb=32.0, group=None):
    """Make ``model.train_step`` data-parallel over the default (or given) process group."""
    store = model.store
    table = [(v.name, v.offset, -(-v.numel // 64) * 64) for v in store.vars.values()]
    red = GradientAllReducer(store.flat["grad"], table, bucket_mb, group, {v.name for v in store.trainable_vars()})
    store.grad_ready_hook = red.on_grad_ready
    model._dp = red
    model.world_size = red.world
    # replicas must start from identical parameters: broadcast rank 0's masters
    if red.world > 1:
        dist.broadcast(store.flat["master"], src=0, group=group)
        store.refresh_half()
    return red
