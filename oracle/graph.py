"""Restatement of the reference's model graphs -- TEST INFRASTRUCTURE ONLY.

The graphs are written once against a tiny backend interface and run on either
  * ``NumpyOps``  -- the NumPy oracle of oracle/tf_ops.py (semantic authority), or
  * ``TorchCpuOps`` -- the same ops expressed with torch CPU kernels (oneDNN convs with
    explicit TF-SAME padding, crop-form transposed convs, matmul, an index-based
    resampler).  It is (a) the autograd cross-check for the conv/FC gradients and
    (b) the multi-threaded "reference CPU path" stand-in timed by bench.py
    (BASELINE.md section 5), since TensorFlow 1.3 cannot run in this image.

Graph sources (paths relative to /root/reference/dyn_mult_view/multi_view_model):
  appearance_flow_model.py:63-66,83-127   AppearanceFlowModel  (M1, M2 base)
  highdim_angle.py:7-10, lowdim_angle.py:7-8   viewpoint-encoder variants (M2)
  appearance_flow_tinghui.py:7-46           lighter Zhou-style net
  main_model.py:57-154                      Base_Prediction_Model (M3, colour+depth)
  multiobject_appflow.py:80-286             MultiObjectAppFlow (M4)
Shape rule: every hard-coded 128-derived size is H/2^k, so the graphs
are parametric in H (multiple of 32); at H=128 they equal the reference literally.
Parameters are a dict keyed by TF variable names ("e0/w", "fc1/Matrix", ...).
"""
import numpy as np

from . import tf_ops as T


# --------------------------------------------------------------------------- #
# backends
# --------------------------------------------------------------------------- #
class NumpyOps:
    name = "numpy"

    def asarray(self, x):
        return np.asarray(x)

    def conv(self, x, w, b, s):
        return T.conv2d_same(x, w, b, s, s)

    def deconv(self, x, w, out_shape, s):
        return T.conv2d_transpose_same(x, w, out_shape, s, s)

    def linear(self, x, m, b):
        return T.linear(x, m, b)

    def lrelu(self, x):
        return T.lrelu(x)

    def relu(self, x):
        return T.relu(x)

    def tanh(self, x):
        return np.tanh(x)

    def concat(self, xs, axis):
        return np.concatenate(xs, axis=axis)

    def flatten_hwc(self, x):
        return x.reshape(x.shape[0], -1)

    def unflatten_hwc(self, x, h, w, c):
        return x.reshape(x.shape[0], h, w, c)

    def split_c(self, x, n):
        return list(np.split(x, n, axis=3))

    def tile_hw(self, v, h, w):
        return np.tile(v[:, None, None, :], (1, h, w, 1))

    def warp_pts(self, flow):
        return T.warp_pts_layer(flow)

    def resample(self, img, warp):
        return T.resampler(img, warp)

    def nhwc(self, x):
        return x

    def from_nhwc(self, x):
        return np.asarray(x)

    def euclidean(self, a, b):
        return T.euclidean_loss(a, b)

    def l1(self, a, b):
        return T.l1_loss(a, b)

    def masked_l2(self, a, b, m):
        d = (a - b) * m
        return np.mean(np.sum(d * d, axis=3, dtype=np.float64), dtype=np.float64)

    def softmax_views(self, logits):  # [V,B,H,W,1] over axis 0
        e = np.exp(logits - logits.max(axis=0, keepdims=True))
        return e / e.sum(axis=0, keepdims=True)


class TorchCpuOps:
    """Activations are carried NCHW-logical / channels_last so oneDNN runs its fast path;
    NHWC numpy arrays go in and come out at the graph boundary."""
    name = "torch-cpu"

    def __init__(self, dtype=None):
        import torch
        self.t = torch
        self.F = torch.nn.functional
        self.dtype = dtype or torch.float32

    def asarray(self, x):
        t = self.t
        return x if isinstance(x, t.Tensor) else t.as_tensor(np.asarray(x), dtype=self.dtype)

    def from_nhwc(self, x):
        x = self.asarray(x)
        return x.permute(0, 3, 1, 2).contiguous(memory_format=self.t.channels_last)

    def nhwc(self, x):
        return x.permute(0, 2, 3, 1)

    def conv(self, x, w, b, s):
        kh, kw = w.shape[0], w.shape[1]
        H, W = x.shape[2], x.shape[3]
        _, pt, pb = T.same_pad(H, kh, s)
        _, pl, pr = T.same_pad(W, kw, s)
        xp = self.F.pad(x, (pl, pr, pt, pb))
        return self.F.conv2d(xp, w.permute(3, 2, 0, 1), b, stride=s)

    def deconv(self, x, w, out_shape, s):
        # TF conv2d_transpose(SAME) == full transposed conv cropped [before : before+out]
        kh, kw = w.shape[0], w.shape[1]
        Ho, Wo = out_shape[1], out_shape[2]
        _, pt, _ = T.same_pad(Ho, kh, s)
        _, pl, _ = T.same_pad(Wo, kw, s)
        full = self.F.conv_transpose2d(x, w.permute(3, 2, 0, 1), None, stride=s)
        need_h, need_w = pt + Ho, pl + Wo
        if full.shape[2] < need_h or full.shape[3] < need_w:
            full = self.F.pad(full, (0, max(0, need_w - full.shape[3]), 0, max(0, need_h - full.shape[2])))
        return full[:, :, pt:pt + Ho, pl:pl + Wo]

    def linear(self, x, m, b):
        return x @ m + b

    def lrelu(self, x):
        return 0.6 * x + 0.4 * x.abs()

    def relu(self, x):
        return 0.5 * x + 0.5 * x.abs()

    def tanh(self, x):
        return self.t.tanh(x)

    def concat(self, xs, axis):
        if xs[0].dim() == 4:
            axis = {0: 0, 1: 2, 2: 3, 3: 1}[axis]
        return self.t.cat(xs, dim=axis)

    def flatten_hwc(self, x):
        return x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)

    def unflatten_hwc(self, x, h, w, c):
        return x.reshape(x.shape[0], h, w, c).permute(0, 3, 1, 2)

    def split_c(self, x, n):
        return list(self.t.chunk(x, n, dim=1))

    def tile_hw(self, v, h, w):
        return v[:, :, None, None].expand(-1, -1, h, w)

    def warp_pts(self, flow):                      # flow NCHW [B,2,H,W] -> NHWC warp [B,H,W,2]
        f = flow.permute(0, 2, 3, 1)
        B, H, W, _ = f.shape
        return f + self.asarray(T.coords(H, W, B)).to(f.dtype)

    def resample(self, img, warp):
        """TF resampler rule with torch indexing; img NCHW, warp NHWC [B,H,W,2] -> NCHW."""
        t = self.t
        data = img.permute(0, 2, 3, 1)              # NHWC view
        B, H, W, C = data.shape
        x, y = warp[..., 0], warp[..., 1]
        valid = (x > -1) & (y > -1) & (x < W) & (y < H)
        xs = t.where(valid, x, t.zeros_like(x))
        ys = t.where(valid, y, t.zeros_like(y))
        fx = t.floor(xs.detach()).long()
        fy = t.floor(ys.detach()).long()
        cx, cy = fx + 1, fy + 1
        dx = cx.to(x.dtype) - xs
        dy = cy.to(y.dtype) - ys
        bidx = t.arange(B).reshape(B, 1, 1).expand_as(fx)

        def P(u, v):
            ok = (u >= 0) & (u <= W - 1) & (v >= 0) & (v <= H - 1)
            vals = data[bidx, v.clamp(0, H - 1), u.clamp(0, W - 1)]
            return vals * ok[..., None].to(vals.dtype)

        dxe, dye = dx[..., None], dy[..., None]
        out = (dxe * dye) * P(fx, fy) + ((1 - dxe) * (1 - dye)) * P(cx, cy) \
            + (dxe * (1 - dye)) * P(fx, cy) + ((1 - dxe) * dye) * P(cx, fy)
        out = out * valid[..., None].to(out.dtype)
        return out.permute(0, 3, 1, 2)

    def euclidean(self, a, b):
        d = a - b
        return (d * d).sum(dim=1).mean()

    def l1(self, a, b):
        return (a - b).abs().sum(dim=1).mean()

    def masked_l2(self, a, b, m):
        d = (a - b) * m
        return (d * d).sum(dim=1).mean()

    def softmax_views(self, logits):
        return self.t.softmax(logits, dim=0)


def _bf16_round_fns(torch):
    """(round_fwd, round_bwd, round_both): straight-through bf16 rounding points for Bf16TorchCpuOps."""
    def r(x):
        return x.to(torch.bfloat16).to(torch.float32)

    class RoundF(torch.autograd.Function):       # value is stored as bf16; gradient passes untouched
        @staticmethod
        def forward(ctx, x):
            return r(x)

        @staticmethod
        def backward(ctx, g):
            return g

    class RoundB(torch.autograd.Function):       # value untouched; the gradient arriving here is stored as bf16
        @staticmethod
        def forward(ctx, x):
            return x.view_as(x)

        @staticmethod
        def backward(ctx, g):
            return r(g)

    class RoundFB(torch.autograd.Function):      # both the value and its gradient are stored as bf16
        @staticmethod
        def forward(ctx, x):
            return r(x)

        @staticmethod
        def backward(ctx, g):
            return r(g)

    return RoundF.apply, RoundB.apply, RoundFB.apply


class Bf16TorchCpuOps(TorchCpuOps):
    """The torch-CPU backend with bf16 rounding exactly where the CUDA path STORES bf16, fp32
    arithmetic everywhere else -- so that a comparison against it isolates kernel error from the quantisation that the
    bf16 data layout itself introduces (against the plain fp32 backends the two are mixed).  Rounding points:
      * the bf16 compute copy of every conv / deconv / FC weight (biases stay fp32 masters);
      * the input of every conv / deconv / FC (network inputs are cast: thin_s2d_prep, dmv_cast_f32_to_bf16; hidden
        activations are already bf16, the rounding is idempotent);
      * the output of every lrelu / relu epilogue (stored bf16); tanh / linear heads write fp32 and are NOT rounded;
      * backward: the gradient leaving the activation derivative, dPre = dX_next * act'(Y), is written as bf16 -- by the
        consuming layer's input-gradient epilogue, which applies act'(its input) to the fp32 accumulator before its one
        rounding (include/dmv3d.h, y_in / act_in), or by thin_s2d_prep for the fp32 heads; both dgrad and wgrad consume
        that rounded dPre.  (Where a torch op sits between two layers -- concat, channel split -- dX is rounded once more
        before the factor is applied; the difference is below the test tolerances and not modelled.)
    Weight gradients, the sampler, the losses and Adam are fp32 on both sides."""
    name = "torch-cpu"

    def __init__(self):
        super().__init__()
        self.rf, self.rb, self.rfb = _bf16_round_fns(self.t)

    def conv(self, x, w, b, s):
        return self.rb(super().conv(self.rf(x), self.rf(w), b, s))

    def deconv(self, x, w, out_shape, s):
        return self.rb(super().deconv(self.rf(x), self.rf(w), out_shape, s))

    def linear(self, x, m, b):
        return self.rb(super().linear(self.rf(x), self.rf(m), b))

    def lrelu(self, x):
        return self.rf(super().lrelu(x))

    def relu(self, x):
        return self.rf(super().relu(x))


# --------------------------------------------------------------------------- #
# parameter shapes (TF variable names) -- shared by oracle users
# --------------------------------------------------------------------------- #
def _conv(shapes, name, k, cin, cout):
    shapes[name + "/w"] = ("conv", (k, k, cin, cout))
    shapes[name + "/b"] = ("zero", (cout,))


def _deconv(shapes, name, k, cout, cin, stride=2):
    shapes[name + "/w"] = ("deconv%d" % stride, (k, k, cout, cin))


def _fc(shapes, name, kin, nout):
    shapes[name + "/Matrix"] = ("fc", (kin, nout))
    shapes[name + "/b"] = ("zero", (nout,))


def appflow_param_shapes(H, V, kind="base"):
    """Variable shapes of AppearanceFlowModel and its viewpoint variants at side H."""
    s = {}
    h5 = H // 32
    if kind == "tinghui":
        for name, cin, cout in [("e0", 3, 16), ("e1", 16, 32), ("e2", 32, 64), ("e3", 64, 128), ("e4", 128, 256)]:
            _conv(s, name, 3, cin, cout)
        _fc(s, "e_fc0", h5 * h5 * 256, 2048)
        _fc(s, "e_fc1", 2048, 2048)
        _fc(s, "a0", V, 64); _fc(s, "a1", 64, 64); _fc(s, "a2", 64, 64)
        _fc(s, "a3", 2048 + 64, 2048)
        _fc(s, "a4", 2048, (H // 16) ** 2 * 32)
        _deconv(s, "d3", 3, 128, 32); _deconv(s, "d2", 3, 64, 128)
        _deconv(s, "d1", 3, 32, 64); _deconv(s, "d0", 3, 16, 32)
        _deconv(s, "flow_field", 3, 2, 16, stride=1)
        return s
    for name, k, cin, cout in [("e0", 5, 3, 32), ("e0_0", 5, 32, 32), ("e1", 5, 32, 32), ("e1_0", 5, 32, 32),
                               ("e2", 5, 32, 64), ("e2_0", 5, 64, 64), ("e3", 3, 64, 128), ("e3_0", 3, 128, 128),
                               ("e4", 3, 128, 256), ("e4_0", 3, 256, 256)]:
        _conv(s, name, k, cin, cout)
    _fc(s, "fc1", h5 * h5 * 256, 4096)
    if kind == "base":
        _fc(s, "a0", V, 64); _fc(s, "a1", 64, 64); _fc(s, "a2", 64, 64); A = 64
    elif kind == "highdim":          # highdim_angle.py:8-10 : a0, a1 are dead variables
        _fc(s, "a0", V, 19); _fc(s, "a1", V, 128); _fc(s, "a2", V, 256); A = 256
    elif kind == "lowdim":           # lowdim_angle.py:8
        _fc(s, "a0", V, 10); A = 10
    else:
        raise ValueError(kind)
    _fc(s, "a3", 4096 + A, 4096)
    _fc(s, "a4", 4096, 4096)
    _fc(s, "a5", 4096, h5 * h5 * 256)
    _deconv(s, "d4", 3, 128, 256); _conv(s, "d4_0", 3, 128, 128)
    _deconv(s, "d3", 3, 64, 128); _conv(s, "d3_0", 5, 64, 64)
    _deconv(s, "d2", 5, 32, 64); _conv(s, "d2_0", 5, 32, 64)
    _deconv(s, "d1", 5, 32, 64); _conv(s, "d1_0", 5, 32, 32)
    _deconv(s, "flow_field", 5, 2, 32)
    return s


def init_params(shapes, seed=0):
    """Reference initialisers (tf_utils.py:54-98) drawn with NumPy for oracle-only use."""
    rng = np.random.default_rng(seed)
    p = {}
    for name, (kind, shp) in shapes.items():
        if kind == "zero":
            p[name] = np.zeros(shp, np.float32)
        elif kind == "fc":
            p[name] = (rng.standard_normal(shp) * T.linear_stddev(shp[0])).astype(np.float32)
        elif kind == "conv":
            sd = T.conv_stddev(shp[0], shp[1], shp[2])
            v = rng.standard_normal(shp)
            bad = np.abs(v) > 2
            while bad.any():                       # truncated_normal: re-draw beyond 2 sigma
                v[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(v) > 2
            p[name] = (v * sd).astype(np.float32)
        elif kind.startswith("deconv"):
            st = int(kind[6:])
            p[name] = (rng.standard_normal(shp) * T.deconv_stddev(shp[0], shp[1], shp[3], st, st)).astype(np.float32)
        else:
            raise ValueError(kind)
    return p


# --------------------------------------------------------------------------- #
# teacher forcing: the checked implementation's own stored activations as layer inputs
# --------------------------------------------------------------------------- #
_FORCE = [None]
_FREE = [None]
_FORCE_GRAD = [None]
_FREE_GRAD = [None]


class forcing:
    """Context for the torch backends: ``with forcing(record, grad_record) as f:`` runs a graph in which the OUTPUT of
    every layer named in ``record`` (name -> NHWC / [B,N] activation as the implementation under test stored it) replaces
    the oracle's own, value and activation slope alike, while the autograd path through the oracle's arithmetic stays;
    with ``grad_record`` (name -> the gradient the implementation received for that activation) the incoming gradient
    of every such layer is replaced likewise during backward.

    Why: bf16 rounding is chaotic.  Two correct implementations whose fp32 sums differ in the last bit round a few
    activations to different bf16 neighbours; the next layer amplifies a perturbation d to sqrt(d * ulp), so after a few
    layers the two chains differ by the full bf16 quantisation noise (~sqrt(L) * 2^-9), and leaky-relu units near zero
    take different slopes, which changes bottleneck gradients by 10 % and more -- between ANY two bf16 chains, and
    between a bf16 chain and the fp32 reference.  With forcing, every layer is checked on identical inputs:
    ``f.free[name]`` is the oracle's own output of the layer given the implementation's input (forward check),
    ``f.free_grad[name]`` the oracle's own gradient for that activation given the implementation's downstream gradient
    (input-gradient check of the NEXT layer), and every parameter gradient is computed from identical activations,
    slopes and output gradients (weight-gradient check).  A wrong kernel at layer L still fails: its output differs from
    ``free[L]``, its gradients from the oracle's."""

    def __init__(self, record, grad_record=None):
        """grad_record: name -> (kind, tensor); kind "out" = gradient w.r.t. the layer's output, "pre" = w.r.t. its
        pre-activation (what an implementation with the activation derivative fused into the consumer's dgrad has)."""
        self.record, self.grad_record = record, grad_record
        self.free, self.free_grad = {}, {}

    def __enter__(self):
        _FORCE[0], _FREE[0], _FORCE_GRAD[0], _FREE_GRAD[0] = self.record, self.free, self.grad_record, self.free_grad
        return self

    def __exit__(self, *a):
        _FORCE[0] = _FREE[0] = _FORCE_GRAD[0] = _FREE_GRAD[0] = None


def _force_grad(ops, name, y, kind):
    gr, free = _FORCE_GRAD[0], _FREE_GRAD[0]
    if gr is None or name not in gr or gr[name][0] != kind:
        return y
    t = ops.t
    g_forced = t.as_tensor(gr[name][1]).detach().to(t.float32)
    g_forced = ops.from_nhwc(g_forced) if g_forced.dim() == 4 else g_forced

    class ForceGrad(t.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return x.view_as(x)

        @staticmethod
        def backward(ctx, g):
            free[name] = g.detach()
            return g_forced.reshape(g.shape)

    return ForceGrad.apply(y)


def _finish(ops, name, pre, act, acts=None):
    """Activation epilogue of layer ``name`` (act in 'lrelu' | 'relu' | 'tanh' | None) + the forcing hook."""
    fn = {"lrelu": ops.lrelu, "relu": ops.relu, "tanh": ops.tanh, None: (lambda v: v)}[act]
    y = fn(pre)
    if acts is not None:
        acts[name] = y
    f = _FORCE[0]
    if f is not None and name in f and act != "tanh":
        t = ops.t
        _FREE[0][name] = y.detach()
        yc = t.as_tensor(f[name]).detach().to(t.float32)
        yc = ops.from_nhwc(yc) if yc.dim() == 4 else yc
        if act == "lrelu":                       # invert the activation: the slope then follows the FORCED output's sign
            pc = t.where(yc > 0, yc, yc * 5.0)
        elif act == "relu":
            pc = t.where(yc > 0, yc, -t.ones_like(yc))
        else:
            pc = yc
        y = _force_grad(ops, name, fn(_force_grad(ops, name, pre + (pc - pre).detach(), "pre")), "out")
    return y


# --------------------------------------------------------------------------- #
# graphs
# --------------------------------------------------------------------------- #
def _decode_angle(ops, P, disp, kind):
    if kind in ("base", "tinghui"):          # appearance_flow_model.py:63-66 / tinghui :7-11
        a0 = _finish(ops, "a0", ops.linear(disp, P["a0/Matrix"], P["a0/b"]), "lrelu")
        a1 = _finish(ops, "a1", ops.linear(a0, P["a1/Matrix"], P["a1/b"]), "lrelu")
        return _finish(ops, "a2", ops.linear(a1, P["a2/Matrix"], P["a2/b"]), "lrelu")
    if kind == "highdim":                    # highdim_angle.py:8-10 (a0, a1 outputs unused)
        return _finish(ops, "a2", ops.linear(disp, P["a2/Matrix"], P["a2/b"]), "lrelu")
    if kind == "lowdim":                     # lowdim_angle.py:8
        return _finish(ops, "a0", ops.linear(disp, P["a0/Matrix"], P["a0/b"]), "lrelu")
    raise ValueError(kind)


def appearance_flow_forward(ops, params, image0, disp, kind="base", keep=False):
    """appearance_flow_model.py:83-127 (kind base/highdim/lowdim) or
    appearance_flow_tinghui.py:13-46 (kind tinghui).  image0 NHWC [B,H,H,3], disp [B,V].
    Returns dict with flow_field, warp_pts, gen (NHWC) and, if keep, every activation."""
    P = {k: ops.asarray(v) for k, v in params.items()}
    x = ops.from_nhwc(image0)
    disp = ops.asarray(disp)
    B, H = image0.shape[0], image0.shape[1]
    acts = {}

    def C(name, inp, s, act="lrelu"):
        return _finish(ops, name, ops.conv(inp, P[name + "/w"], P[name + "/b"], s), act, acts)

    def D(name, inp, side, cout, s=2, act="lrelu"):
        return _finish(ops, name, ops.deconv(inp, P[name + "/w"], (B, side, side, cout), s), act, acts)

    def FC(name, inp, act="lrelu"):
        return _finish(ops, name, ops.linear(inp, P[name + "/Matrix"], P[name + "/b"]), act, acts)

    h5 = H // 32
    if kind == "tinghui":
        r = "relu"
        e = C("e0", x, 2, r); e = C("e1", e, 2, r); e = C("e2", e, 2, r); e = C("e3", e, 2, r); e = C("e4", e, 2, r)
        f = FC("e_fc0", ops.flatten_hwc(e), r)
        f = FC("e_fc1", f, r)
        ang = _decode_angle(ops, P, disp, kind)
        j = FC("a3", ops.concat([f, ang], 1), r)
        j = FC("a4", j, r)
        d = ops.unflatten_hwc(j, H // 16, H // 16, 32)
        d = D("d3", d, H // 8, 128, 2, r); d = D("d2", d, H // 4, 64, 2, r)
        d = D("d1", d, H // 2, 32, 2, r); d = D("d0", d, H, 16, 2, r)
        flow = D("flow_field", d, H, 2, 1, None)
    else:
        e = C("e0", x, 2); e = C("e0_0", e, 1); e = C("e1", e, 2); e = C("e1_0", e, 1)
        e = C("e2", e, 2); e = C("e2_0", e, 1); e = C("e3", e, 2); e = C("e3_0", e, 1)
        e = C("e4", e, 2); e = C("e4_0", e, 1)
        e5 = FC("fc1", ops.flatten_hwc(e))
        ang = _decode_angle(ops, P, disp, kind)
        acts["angle"] = ang
        j = FC("a3", ops.concat([e5, ang], 1)); j = FC("a4", j); j = FC("a5", j)
        d = ops.unflatten_hwc(j, h5, h5, 256)
        d = D("d4", d, 2 * h5, 128); d = C("d4_0", d, 1)
        d = D("d3", d, 4 * h5, 64); d = C("d3_0", d, 1)
        d = D("d2", d, 8 * h5, 32); d = C("d2_0", d, 1)
        d = D("d1", d, 16 * h5, 32); d = C("d1_0", d, 1)
        flow = D("flow_field", d, H, 2, 2, None)
    warp = ops.warp_pts(flow)
    gen = ops.resample(x, warp)
    out = {"flow_field": ops.nhwc(flow), "warp_pts": warp, "gen": ops.nhwc(gen)}
    if keep:
        out["acts"] = {k: (ops.nhwc(v) if getattr(v, "ndim", 0) == 4 else v) for k, v in acts.items()}
    return out


def appearance_flow_loss(ops, out, image1, mode="l2"):
    """appearance_flow_model.py:73 (euclidean) or the north-star's L1 (tf_utils.py:22)."""
    tgt = ops.asarray(image1)
    gen = out["gen"]
    if ops.name == "torch-cpu":
        gen = gen.permute(0, 3, 1, 2)
        tgt = tgt.permute(0, 3, 1, 2)
    return ops.euclidean(gen, tgt) if mode == "l2" else ops.l1(gen, tgt)


# --- M3 : main_model.py ------------------------------------------------------ #
def _pre_encoder_shapes(s, scope, cin):
    for name, ci, co in [("e0", cin, 32), ("e0_0", 32, 32), ("e1", 32, 32), ("e1_0", 32, 32), ("e2", 32, 64)]:
        _conv(s, scope + "/" + name, 5, ci, co)


def _decoder_shapes(s, scope, cout):
    _deconv(s, scope + "/d2", 5, 32, 64); _conv(s, scope + "/d2_0", 5, 32, 64)
    _deconv(s, scope + "/d1", 5, 32, 64); _conv(s, scope + "/d1_0", 5, 32, 32)
    _deconv(s, scope + "/d0", 5, cout, 32)


def _trunk_shapes(s, H, V, n_in, num_decode, fully_conv=False):
    h5 = H // 32
    _conv(s, "e2_0", 5, 64 * n_in, 64); _conv(s, "e3", 3, 64, 128); _conv(s, "e3_0", 3, 128, 128)
    _conv(s, "e4", 3, 128, 256); _conv(s, "e4_0", 3, 256, 256)
    _fc(s, "a0", V, 64); _fc(s, "a1", 64, 64); _fc(s, "a2", 64, 64)
    if fully_conv:
        _conv(s, "e4_1", 3, 256 + 64, 256); _conv(s, "e4_2", 3, 256, 256)
    else:
        _fc(s, "fc1", h5 * h5 * 256, 4096)
        _fc(s, "a3", 4096 + 64, 4096); _fc(s, "a4", 4096, 4096); _fc(s, "a5", 4096, h5 * h5 * 256)
    _deconv(s, "d4", 3, 128, 256); _conv(s, "d4_0", 3, 128, 128)
    _deconv(s, "d3", 3, 64, 128); _conv(s, "d3_0", 5, 64, 64 * num_decode)


def colordepth_param_shapes(H, V, conf):
    """Base_Prediction_Model (main_model.py:83-142).  ``head`` mode 'tanh' is the
    reference; 'flow' replaces each tanh head by a 2-ch flow head + sampler."""
    s = {}
    n = 0
    if "use_color" in conf:
        _pre_encoder_shapes(s, "pre_image0", 3); n += 1
    if "use_depth" in conf:
        _pre_encoder_shapes(s, "pre_dimage0", 1); n += 1
    _trunk_shapes(s, H, V, n, n)
    flow = conf.get("head", "tanh") == "flow"
    if "use_color" in conf:
        _decoder_shapes(s, "dec_image1", 2 if flow else 3)
    if "use_depth" in conf:
        _decoder_shapes(s, "dec_dimage1", 2 if flow else 1)
    return s


def _pre_encode(ops, P, x, scope):
    for name, st in [("e0", 2), ("e0_0", 1), ("e1", 2), ("e1_0", 1), ("e2", 2)]:
        x = _finish(ops, "%s/%s" % (scope, name), ops.conv(x, P["%s/%s/w" % (scope, name)], P["%s/%s/b" % (scope, name)], st), "lrelu")
    return x


def _trunk(ops, P, comb, disp, B, H, fully_conv=False):
    h5 = H // 32
    c = lambda n, x, s: _finish(ops, n, ops.conv(x, P[n + "/w"], P[n + "/b"], s), "lrelu")
    f = lambda n, x: _finish(ops, n, ops.linear(x, P[n + "/Matrix"], P[n + "/b"]), "lrelu")
    e = c("e2_0", comb, 1); e = c("e3", e, 2); e = c("e3_0", e, 1); e = c("e4", e, 2); e = c("e4_0", e, 1)
    a2 = f("a2", f("a1", f("a0", disp)))
    if fully_conv:                      # multiobject_appflow.py:147-153
        sm = ops.tile_hw(a2, h5, h5)
        e = c("e4_1", ops.concat([e, sm], 3), 1)
        a5r = c("e4_2", e, 1)
    else:
        e5 = f("fc1", ops.flatten_hwc(e))
        j = f("a5", f("a4", f("a3", ops.concat([e5, a2], 1))))
        a5r = ops.unflatten_hwc(j, h5, h5, 256)
    d = _finish(ops, "d4", ops.deconv(a5r, P["d4/w"], (B, 2 * h5, 2 * h5, 128), 2), "lrelu")
    d = c("d4_0", d, 1)
    d = _finish(ops, "d3", ops.deconv(d, P["d3/w"], (B, 4 * h5, 4 * h5, 64), 2), "lrelu")
    return c("d3_0", d, 1)


def _decode(ops, P, x, scope, B, H, cout):
    h5 = H // 32
    d = _finish(ops, scope + "/d2", ops.deconv(x, P[scope + "/d2/w"], (B, 8 * h5, 8 * h5, 32), 2), "lrelu")
    d = _finish(ops, scope + "/d2_0", ops.conv(d, P[scope + "/d2_0/w"], P[scope + "/d2_0/b"], 1), "lrelu")
    d = _finish(ops, scope + "/d1", ops.deconv(d, P[scope + "/d1/w"], (B, 16 * h5, 16 * h5, 32), 2), "lrelu")
    d = _finish(ops, scope + "/d1_0", ops.conv(d, P[scope + "/d1_0/w"], P[scope + "/d1_0/b"], 1), "lrelu")
    return ops.deconv(d, P[scope + "/d0/w"], (B, H, H, cout), 2)


def colordepth_forward(ops, params, conf, image0, dimage0, disp):
    """main_model.py:83-142.  split_list.pop() hands out channel groups from the END."""
    P = {k: ops.asarray(v) for k, v in params.items()}
    B, H = image0.shape[0], image0.shape[1]
    disp = ops.asarray(disp)
    feats, srcs = [], {}
    if "use_color" in conf:
        srcs["color"] = ops.from_nhwc(image0)
        feats.append(_pre_encode(ops, P, srcs["color"], "pre_image0"))
    if "use_depth" in conf:
        srcs["depth"] = ops.from_nhwc(dimage0)
        feats.append(_pre_encode(ops, P, srcs["depth"], "pre_dimage0"))
    d3_0 = _trunk(ops, P, ops.concat(feats, 3), disp, B, H)
    split = ops.split_c(d3_0, len(feats))
    flow = conf.get("head", "tanh") == "flow"
    out = {}
    for key, scope, name, cout in [("color", "dec_image1", "gen_image1", 3), ("depth", "dec_dimage1", "gen_dimage1", 1)]:
        if ("use_" + key) not in conf:
            continue
        pre = _decode(ops, P, split.pop(), scope, B, H, 2 if flow else cout)
        if flow:
            out[name] = ops.nhwc(ops.resample(srcs[key], ops.warp_pts(pre)))
            out[name + "_flow"] = ops.nhwc(pre)
        else:
            out[name] = ops.nhwc(ops.tanh(pre))
    return out


def colordepth_loss(ops, out, conf, image1, dimage1, mode="l2"):
    """main_model.py:144-154: L(color) + depth_lr_factor * L(depth); mode 'l1' is the
    BASELINE config-4 "per-channel L1" (tf_utils.py:22, weights as mv3d/nobg_dm.py:91-92)."""
    L = ops.euclidean if mode == "l2" else ops.l1
    tot = 0.0
    perm = (lambda x: x.permute(0, 3, 1, 2)) if ops.name == "torch-cpu" else (lambda x: x)
    if "use_color" in conf:
        tot = tot + L(perm(out["gen_image1"]), perm(ops.asarray(image1)))
    if "use_depth" in conf:
        tot = tot + L(perm(out["gen_dimage1"]), perm(ops.asarray(dimage1))) * conf["depth_lr_factor"]
    return tot


# --- 8(f)-3 : multi-view confidence-weighted fusion (NOT in the reference) ----- #
def fuse_views(ops, gens, logits):
    """out = sum_v softmax_v(c)_v * gen_v ; gens [V,B,H,W,C], logits [V,B,H,W,1].
    Definition adopted in SURVEY 8(f)-3 (after Zhou et al. 2016) -- parity unpinned."""
    w = ops.softmax_views(logits)
    return (w * gens).sum(0)


# --- M4 : multiobject_appflow.py ------------------------------------------------ #
_MO_INPUTS = [("use_color", "image0", "pre_image0_f", 3), ("use_depth", "depth0", "pre_dimage0_f", 1),
              (None, "image0_mask0", "pre_mask0_ob0", 1), (None, "image0_mask1", "pre_mask0_ob1", 1)]


def _mo_heads(conf):
    """Decoder heads in the order multiobject_appflow.py:189-218 pops them: (attribute, scope, kind)."""
    heads = []
    if "use_color" in conf:
        if "combination_image" in conf:
            heads.append(("gen_image1", "dec_image1", "flow"))
        if "gen_sep_images" in conf:
            heads += [("gen_image1_only0", "dec_image1_only0", "flow"), ("gen_image1_only1", "dec_image1_only1", "flow")]
    if "use_depth" in conf:
        if "combination_image" in conf:
            heads.append(("gen_depth1", "dec_dimage1_f", "tanh"))
        if "gen_sep_images" in conf:
            heads += [("gen_depth1_only0", "dec_depth1_only0", "tanh"), ("gen_depth1_only1", "dec_depth1_only1", "tanh")]
    if "predict_target_masks" in conf:
        heads += [("gen_image1_mask0", "dec_image1_mask0", "tanh"), ("gen_image1_mask1", "dec_image1_mask1", "tanh")]
    return heads


def multiobject_param_shapes(H, V, conf):
    """MultiObjectAppFlow (multiobject_appflow.py:80-221): variables in creation order."""
    s = {}
    n_in = 0
    for key, _, scope, cin in _MO_INPUTS:
        if key is None or key in conf:
            _pre_encoder_shapes(s, scope, cin); n_in += 1
    heads = _mo_heads(conf)
    _trunk_shapes(s, H, V, n_in, len(heads), fully_conv="fully_conv" in conf)
    for _, scope, kind in heads:
        _decoder_shapes(s, scope, 2 if kind == "flow" else 1)
    return s


def multiobject_forward(ops, params, conf, batch):
    """multiobject_appflow.py:123-221.  batch: dict of NHWC arrays under the reference's attribute names.
    Every flow head samples image0 (:193-198); split_list.pop() hands out channel groups from the END."""
    P = {k: ops.asarray(v) for k, v in params.items()}
    B, H = batch["image0"].shape[0], batch["image0"].shape[1]
    feats = []
    for key, attr, scope, _ in _MO_INPUTS:
        if key is None or key in conf:
            feats.append(_pre_encode(ops, P, ops.from_nhwc(batch[attr]), scope))
    heads = _mo_heads(conf)
    d3_0 = _trunk(ops, P, ops.concat(feats, 3), ops.asarray(batch["displacement"]), B, H, fully_conv="fully_conv" in conf)
    split = ops.split_c(d3_0, len(heads))
    src = ops.from_nhwc(batch["image0"])
    out = {}
    for attr, scope, kind in heads:
        pre = _decode(ops, P, split.pop(), scope, B, H, 2 if kind == "flow" else 1)
        out[attr] = ops.nhwc(ops.resample(src, ops.warp_pts(pre))) if kind == "flow" else ops.nhwc(ops.tanh(pre))
    return out


def multiobject_loss(ops, out, conf, batch):
    """multiobject_appflow.py:223-283."""
    perm = (lambda x: x.permute(0, 3, 1, 2)) if ops.name == "torch-cpu" else (lambda x: x)
    A = lambda k: perm(ops.asarray(batch[k]))
    G = lambda k: perm(out[k])
    tot = 0.0

    def pair(gen, tgt, mask, factor):
        if "masked_image_loss" in conf:
            return ops.masked_l2(G(gen), A(tgt), A(mask)) * factor
        return ops.euclidean(G(gen), A(tgt)) * factor

    if "use_color" in conf:
        if "combination_image" in conf:
            tot = tot + ops.euclidean(G("gen_image1"), A("image1"))
        if "gen_sep_images" in conf:
            tot = tot + pair("gen_image1_only0", "image1_only0", "image1_mask0", 1.0)
            tot = tot + pair("gen_image1_only1", "image1_only1", "image1_mask1", 1.0)
    if "use_depth" in conf:
        f = conf["use_depth"]
        if "combination_image" in conf:
            tot = tot + ops.euclidean(G("gen_depth1"), A("depth1"))          # :260 -- no depth factor on the combined image
        if "gen_sep_images" in conf:
            tot = tot + pair("gen_depth1_only0", "depth1_only0", "image1_mask0", f)
            tot = tot + pair("gen_depth1_only1", "depth1_only1", "image1_mask1", f)
    if "predict_target_masks" in conf:
        f = conf["predict_target_masks"]
        tot = tot + ops.euclidean(G("gen_image1_mask0"), A("image1_mask0")) * f
        tot = tot + ops.euclidean(G("gen_image1_mask1"), A("image1_mask1")) * f
    return tot


# --- config 5 : multi-view appearance flow with confidence fusion (NOT in the reference) ---------- #
_MV_CONF_KEYS = ("use_color", "use_depth", "fully_conv")


def _mv_conf(conf):
    c = {k: conf[k] for k in _MV_CONF_KEYS if k in conf}
    c.setdefault("use_color", "")
    return c


def multiview_param_shapes(H, V, conf=None):
    """SURVEY 8(f)-3 on the multi-object trunk (multiobject_appflow.py:123-187): the pre-encoders of the inputs the conf
    selects (colour, depth, the two object masks), the shared trunk, ONE decoder `dec_image1` whose last layer `d0` has 3
    channels (flow x, flow y, confidence logit).  The same weights serve every source frame."""
    conf = _mv_conf(conf or {})
    s = {}
    n_in = 0
    for key, _, scope, cin in _MO_INPUTS:
        if key is None or key in conf:
            _pre_encoder_shapes(s, scope, cin); n_in += 1
    _trunk_shapes(s, H, V, n_in, 1, fully_conv="fully_conv" in conf)
    _decoder_shapes(s, "dec_image1", 3)
    return s


def multiview_forward(ops, params, conf, batch):
    """batch: image0 [Vw,B,H,H,3] source frames, (depth0,) image0_mask0/1 [Vw,B,H,H,1], displacement [Vw,B,V] (each
    frame's viewpoint change to the target).  Frame v runs the multi-object trunk (shared weights) and one 3-channel
    head: flow_v = head[...,0:2], logit_v = head[...,2]; gen_v = warp(image0_v, flow_v);
    fused = sum_v softmax_v(logit)_v * gen_v   (after Zhou et al. 2016).
    Every layer acts per sample, so the Vw frames of the B samples are run as one batch of Vw*B."""
    conf = _mv_conf(conf or {})
    P = {k: ops.asarray(v) for k, v in params.items()}
    Vw, B, H = batch["image0"].shape[0], batch["image0"].shape[1], batch["image0"].shape[2]
    N = Vw * B
    flat = lambda k: np.asarray(batch[k]).reshape((N,) + tuple(batch[k].shape[2:]))
    feats = []
    for key, attr, scope, _ in _MO_INPUTS:
        if key is None or key in conf:
            feats.append(_pre_encode(ops, P, ops.from_nhwc(flat(attr)), scope))
    d3_0 = _trunk(ops, P, ops.concat(feats, 3), ops.asarray(flat("displacement")), N, H, fully_conv="fully_conv" in conf)
    head = ops.nhwc(_decode(ops, P, d3_0, "dec_image1", N, H, 3))                  # NHWC [Vw*B,H,H,3]
    flow, logit = head[..., 0:2], head[..., 2:3]
    gen = ops.nhwc(ops.resample(ops.from_nhwc(flat("image0")), flow + ops.asarray(T.coords(H, H, N))))
    gens, logits = gen.reshape(Vw, B, H, H, 3), logit.reshape(Vw, B, H, H, 1)
    return {"gens": gens, "logits": logits[..., 0], "flows": flow.reshape(Vw, B, H, H, 2), "fused": fuse_views(ops, gens, logits)}


def multiview_loss(ops, out, image1, mode="l2"):
    tgt = ops.asarray(image1)
    fused = out["fused"]
    if ops.name == "torch-cpu":
        fused, tgt = fused.permute(0, 3, 1, 2), tgt.permute(0, 3, 1, 2)
    return ops.euclidean(fused, tgt) if mode == "l2" else ops.l1(fused, tgt)
