"""NumPy restatement of the reference's op helpers -- TEST INFRASTRUCTURE ONLY.

Every function cites the reference line it follows (paths relative to
/root/reference).  Where the reference delegates to TensorFlow 1.3 (un-vendored),
the TF semantics written out in SURVEY.md 8(a) are restated here.
Layout: NHWC, float32 unless ``dtype=np.float64`` is passed for gradient checks.
"""
import math
import numpy as np


# --------------------------------------------------------------------------- #
# activations / losses : dyn_mult_view/mv3d/utils/tf_utils.py:18-33
# --------------------------------------------------------------------------- #
def euclidean_loss(a, b):
    """tf_utils.py:18-19  reduce_mean(reduce_sum((a-b)^2, axis=3))."""
    d = a - b
    return np.mean(np.sum(d * d, axis=3, dtype=np.float64), dtype=np.float64)


def euclidean_loss_grad(a, b):
    """d euclidean_loss / d a  (TF gradient of mean(sum(pow(sub)))))."""
    n = a.shape[0] * a.shape[1] * a.shape[2]
    return (2.0 * (a - b) / n).astype(a.dtype)


def l1_loss(a, b):
    """tf_utils.py:22-23  reduce_mean(reduce_sum(abs(a-b), axis=3))."""
    return np.mean(np.sum(np.abs(a - b), axis=3, dtype=np.float64), dtype=np.float64)


def l1_loss_grad(a, b):
    """TF: grad(abs) = sign, sign(0) = 0."""
    n = a.shape[0] * a.shape[1] * a.shape[2]
    return (np.sign(a - b) / n).astype(a.dtype)


def relu(x):
    """tf_utils.py:25-27  0.5*x + 0.5*abs(x)  (algebraic form kept)."""
    return (0.5 * x + 0.5 * np.abs(x)).astype(x.dtype)


def lrelu(x, leak=0.2):
    """tf_utils.py:29-33  f1*x + f2*abs(x), f1=.5(1+leak), f2=.5(1-leak)."""
    f1 = 0.5 * (1 + leak)
    f2 = 0.5 * (1 - leak)
    return (f1 * x + f2 * np.abs(x)).astype(x.dtype)


def lrelu_grad(x, g, leak=0.2):
    """TF: d/dx (f1 x + f2 |x|) = f1 + f2 sign(x)  (sign(0)=0)."""
    f1 = 0.5 * (1 + leak)
    f2 = 0.5 * (1 - leak)
    return (g * (f1 + f2 * np.sign(x))).astype(x.dtype)


# --------------------------------------------------------------------------- #
# grid : tf_utils.py:35-52
# --------------------------------------------------------------------------- #
def coords(h, w, batch_size, dtype=np.float32):
    """tf_utils.py:44-52.  X,Y = meshgrid(x,y) ('xy' indexing: X[i,j]=j, Y[i,j]=i);
    coords = tile(stack((Y,X),axis=2)) -> channel 0 = ROW index, channel 1 = COLUMN
    index.  The resampler reads channel 0 as x (column): the (Y,X) quirk."""
    y = np.arange(h, dtype=dtype)
    x = np.arange(w, dtype=dtype)
    X, Y = np.meshgrid(x, y)
    c = np.stack((Y, X), axis=2)[None]
    return np.tile(c, (batch_size, 1, 1, 1))


def warp_pts_layer(flow_field):
    """tf_utils.py:35-38  flow + coords(H, W, B)."""
    b, h, w, _ = flow_field.shape
    return (flow_field + coords(h, w, b, flow_field.dtype)).astype(flow_field.dtype)


# --------------------------------------------------------------------------- #
# tf.contrib.resampler (TF 1.3, un-vendored; call site tf_utils.py:40-42)
# --------------------------------------------------------------------------- #
def resampler_indices(data_shape, warp):
    """Corner indices and predicates of SURVEY 8(a) S1 -- the bit-exact targets.

    Returns (fx, fy, cx, cy) int32 and mask uint8 with
      bit0: sample valid  (x > -1 && y > -1 && x < W && y < H)
      bit1: tap (fx,fy) in range   bit2: (cx,cy)   bit3: (fx,cy)   bit4: (cx,fy)
    Tap bits are reported only for valid samples (0 otherwise); indices of invalid
    samples are reported as 0.
    """
    _, H, W, _ = data_shape
    x = warp[..., 0]
    y = warp[..., 1]
    with np.errstate(invalid="ignore"):
        valid = (x > -1) & (y > -1) & (x < W) & (y < H)  # NaN -> False
    xs = np.where(valid, x, 0)
    ys = np.where(valid, y, 0)
    fx = np.floor(xs).astype(np.int32)
    fy = np.floor(ys).astype(np.int32)
    cx = fx + 1
    cy = fy + 1

    def inr(u, v):
        return (u >= 0) & (u <= W - 1) & (v >= 0) & (v <= H - 1)

    mask = valid.astype(np.uint8)
    mask |= (valid & inr(fx, fy)).astype(np.uint8) << 1
    mask |= (valid & inr(cx, cy)).astype(np.uint8) << 2
    mask |= (valid & inr(fx, cy)).astype(np.uint8) << 3
    mask |= (valid & inr(cx, fy)).astype(np.uint8) << 4
    z = np.int32(0)
    return (np.where(valid, fx, z), np.where(valid, fy, z),
            np.where(valid, cx, z), np.where(valid, cy, z), mask)


def _taps(data, warp):
    """Shared by forward and gradient: weights (in warp's dtype) and zero-padded
    tap values P(u,v,c) = data[b,v,u,c] if in range else 0."""
    B, H, W, C = data.shape
    dt = warp.dtype
    fx, fy, cx, cy, mask = resampler_indices(data.shape, warp)
    valid = (mask & 1).astype(bool)
    x = np.where(valid, warp[..., 0], 0).astype(dt)
    y = np.where(valid, warp[..., 1], 0).astype(dt)
    dx = (cx.astype(dt) - x).astype(dt)
    dy = (cy.astype(dt) - y).astype(dt)
    bidx = np.arange(B).reshape((B,) + (1,) * (warp.ndim - 2))

    def P(u, v, bit):
        ok = ((mask >> bit) & 1).astype(bool)
        uu = np.clip(u, 0, W - 1)
        vv = np.clip(v, 0, H - 1)
        vals = data[bidx, vv, uu]  # [..., C]
        return np.where(ok[..., None], vals, 0).astype(data.dtype), ok

    return valid, (fx, fy, cx, cy), dx, dy, P


def resampler(data, warp):
    """Forward of tf.contrib.resampler.resampler(data[B,H,W,C], warp[B,...,2]).
    warp[...,0] = x (column), warp[...,1] = y (row).  SURVEY 8(a) S1; summation order
    fxfy + cxcy + fxcy + cxfy is kept so a non-contracting fp32 kernel is bit-equal."""
    valid, (fx, fy, cx, cy), dx, dy, P = _taps(data, warp)
    one = dx.dtype.type(1)
    dxe = dx[..., None]
    dye = dy[..., None]
    p_ff, _ = P(fx, fy, 1)
    p_cc, _ = P(cx, cy, 2)
    p_fc, _ = P(fx, cy, 3)
    p_cf, _ = P(cx, fy, 4)
    out = (dxe * dye) * p_ff
    out = out + ((one - dxe) * (one - dye)) * p_cc
    out = out + (dxe * (one - dye)) * p_fc
    out = out + ((one - dxe) * dye) * p_cf
    return np.where(valid[..., None], out, 0).astype(data.dtype)


def resampler_grad(data, warp, grad_out):
    """ResamplerGrad(data, warp, grad_output) -> (grad_data, grad_warp).
    SURVEY 8(a) S2.  grad_warp accumulates over channels in order c = 0..C-1;
    grad_data is a scatter-add over the four taps (only where the tap is in range).
    Invalid samples contribute nothing."""
    B, H, W, C = data.shape
    valid, (fx, fy, cx, cy), dx, dy, P = _taps(data, warp)
    dt = data.dtype
    one = dx.dtype.type(1)
    p_ff, ok_ff = P(fx, fy, 1)
    p_cc, ok_cc = P(cx, cy, 2)
    p_fc, ok_fc = P(fx, cy, 3)
    p_cf, ok_cf = P(cx, fy, 4)
    g = np.where(valid[..., None], grad_out, 0).astype(dt)
    dxe = dx[..., None]
    dye = dy[..., None]

    gw = np.zeros(warp.shape, dtype=dt)
    for c in range(C):  # channel order of the TF loop
        gc = g[..., c]
        gw[..., 0] += gc * ((one - dy) * (p_cc[..., c] - p_fc[..., c]) + dy * (p_cf[..., c] - p_ff[..., c]))
        gw[..., 1] += gc * ((one - dx) * (p_cc[..., c] - p_cf[..., c]) + dx * (p_fc[..., c] - p_ff[..., c]))
    gw = np.where(valid[..., None], gw, 0).astype(dt)

    gd = np.zeros((B, H, W, C), dtype=np.float64)
    bidx = np.broadcast_to(np.arange(B).reshape((B,) + (1,) * (warp.ndim - 2)), fx.shape)

    def scatter(u, v, ok, wgt):
        sel = ok & valid
        contrib = (g * wgt)[sel]
        np.add.at(gd, (bidx[sel], v[sel], u[sel]), contrib)

    scatter(fx, fy, ok_ff, dxe * dye)
    scatter(cx, cy, ok_cc, (one - dxe) * (one - dye))
    scatter(fx, cy, ok_fc, dxe * (one - dye))
    scatter(cx, fy, ok_cf, (one - dxe) * dye)
    return gd.astype(dt), gw


def resample_layer(src_img, warp_pts):
    """tf_utils.py:40-42."""
    return resampler(src_img, warp_pts)


# --------------------------------------------------------------------------- #
# TF 'SAME' padding, conv2d, conv2d_transpose, matmul : tf_utils.py:54-98
# --------------------------------------------------------------------------- #
def same_pad(in_size, k, s):
    """TF SAME: out = ceil(in/s); total = max((out-1)s + k - in, 0);
    before = total // 2, after = total - before  (extra goes bottom/right)."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    before = total // 2
    return out, before, total - before


def _im2col(x, kh, kw, sh, sw):
    B, H, W, C = x.shape
    Ho, pt, pb = same_pad(H, kh, sh)
    Wo, pl, pr = same_pad(W, kw, sw)
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    cols = np.empty((B, Ho, Wo, kh, kw, C), dtype=x.dtype)
    for r in range(kh):
        for s in range(kw):
            cols[:, :, :, r, s, :] = xp[:, r:r + (Ho - 1) * sh + 1:sh, s:s + (Wo - 1) * sw + 1:sw, :]
    return cols, (Ho, Wo, pt, pb, pl, pr)


def _col2im(cols, x_shape, kh, kw, sh, sw):
    B, H, W, C = x_shape
    Ho, pt, pb = same_pad(H, kh, sh)
    Wo, pl, pr = same_pad(W, kw, sw)
    xp = np.zeros((B, H + pt + pb, W + pl + pr, C), dtype=cols.dtype)
    for r in range(kh):
        for s in range(kw):
            xp[:, r:r + (Ho - 1) * sh + 1:sh, s:s + (Wo - 1) * sw + 1:sw, :] += cols[:, :, :, r, s, :]
    return xp[:, pt:pt + H, pl:pl + W, :]


def conv2d_same(x, w, b, sh, sw):
    """tf_utils.py:81-82  tf.nn.conv2d(x, w[kh,kw,Cin,Cout], [1,sh,sw,1], 'SAME') + b."""
    kh, kw, cin, cout = w.shape
    cols, (Ho, Wo, *_rest) = _im2col(x, kh, kw, sh, sw)
    y = cols.reshape(-1, kh * kw * cin) @ w.reshape(kh * kw * cin, cout)
    y = y.reshape(x.shape[0], Ho, Wo, cout)
    if b is not None:
        y = y + b
    return y.astype(x.dtype)


def conv2d_same_grads(x, w, gy, sh, sw):
    """Conv2DBackpropInput / Conv2DBackpropFilter / bias grad of conv2d_same."""
    kh, kw, cin, cout = w.shape
    cols, _ = _im2col(x, kh, kw, sh, sw)
    g2 = gy.reshape(-1, cout)
    gw = (cols.reshape(-1, kh * kw * cin).T @ g2).reshape(w.shape)
    gcols = (g2 @ w.reshape(kh * kw * cin, cout).T).reshape(cols.shape)
    gx = _col2im(gcols, x.shape, kh, kw, sh, sw)
    gb = g2.sum(axis=0)
    return gx.astype(x.dtype), gw.astype(w.dtype), gb.astype(w.dtype)


def conv2d_transpose_same(x, w, out_shape, sh, sw):
    """tf_utils.py:96-97  tf.nn.conv2d_transpose(x, w[kh,kw,Cout,Cin], out_shape, strides)
    (default padding 'SAME', no bias) == input-gradient of a SAME conv that maps
    out_shape -> x.shape with filter w read as [kh,kw,in=Cout,out=Cin]."""
    kh, kw, cout, cin = w.shape
    B, Ho, Wo, _ = out_shape
    assert x.shape[3] == cin and out_shape[3] == cout
    assert same_pad(Ho, kh, sh)[0] == x.shape[1] and same_pad(Wo, kw, sw)[0] == x.shape[2]
    gcols = (x.reshape(-1, cin) @ w.reshape(kh * kw * cout, cin).T)
    gcols = gcols.reshape(B, x.shape[1], x.shape[2], kh, kw, cout)
    return _col2im(gcols, (B, Ho, Wo, cout), kh, kw, sh, sw).astype(x.dtype)


def conv2d_transpose_same_grads(x, w, gy, sh, sw):
    """Gradients of conv2d_transpose_same wrt x and w (gy has out_shape)."""
    kh, kw, cout, cin = w.shape
    cols, _ = _im2col(gy, kh, kw, sh, sw)               # [B,h,w,kh,kw,cout]
    c2 = cols.reshape(-1, kh * kw * cout)
    gx = (c2 @ w.reshape(kh * kw * cout, cin)).reshape(x.shape)
    gw = (c2.T @ x.reshape(-1, cin)).reshape(w.shape)
    return gx.astype(x.dtype), gw.astype(w.dtype)


def linear(x, matrix, b):
    """tf_utils.py:67  tf.matmul(x, Matrix[K,N]) + b."""
    return (x @ matrix + b).astype(x.dtype)


def linear_grads(x, matrix, gy):
    return (gy @ matrix.T).astype(x.dtype), (x.T @ gy).astype(matrix.dtype), gy.sum(axis=0).astype(matrix.dtype)


# --------------------------------------------------------------------------- #
# initialisers : tf_utils.py:54-98 (stddev rules)
# --------------------------------------------------------------------------- #
def linear_stddev(fan_in):
    return math.sqrt(2.0 / float(fan_in))                        # tf_utils.py:58


def conv_stddev(kh, kw, cin):
    return math.sqrt(2.0 / float(kh * kw * cin))                 # tf_utils.py:73-74


def deconv_stddev(kh, kw, cin, dh, dw):
    return math.sqrt(2.0 / float(kh * kw * cin) * float(dh) * float(dw))   # tf_utils.py:90-92


# --------------------------------------------------------------------------- #
# tf.train.AdamOptimizer (TF 1.3) : appearance_flow_model.py:77
# --------------------------------------------------------------------------- #
def adam_tf_step(theta, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """TF-flavoured Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    m <- b1 m + (1-b1) g ; v <- b2 v + (1-b2) g^2 ; theta <- theta - lr_t m/(sqrt(v)+eps).
    eps is NOT bias-corrected (differs from torch.optim.Adam).  t starts at 1.
    All arithmetic in float32 as ApplyAdam does."""
    f = np.float32
    lr_t = adam_lr_t(lr, t, beta1, beta2)
    # TF's ApplyAdam functor form: m += (g - m)(1-b1); v += (g^2 - v)(1-b2);
    # var -= (m * lr_t) / (sqrt(v) + eps)
    # (T(1) - beta) is evaluated in float32 by TF: 1 - 0.999f = 0.0010000467, not float32(0.001)
    m = (m + (g - m) * (f(1.0) - f(beta1))).astype(np.float32)
    v = (v + (g * g - v) * (f(1.0) - f(beta2))).astype(np.float32)
    theta = (theta - (m * lr_t) / (np.sqrt(v) + f(eps))).astype(np.float32)
    return theta, m, v


def adam_lr_t(lr, t, beta1=0.9, beta2=0.999):
    """Bias-corrected step size of TF Adam.  TF keeps beta1_power / beta2_power as float32
    variables multiplied once per step and evaluates lr*sqrt(1-b2p)/(1-b1p) in float32."""
    f = np.float32
    p1, p2 = f(1.0), f(1.0)
    for _ in range(int(t)):
        p1 = f(p1 * f(beta1))
        p2 = f(p2 * f(beta2))
    return f(f(lr) * np.sqrt(f(1.0) - p2, dtype=np.float32) / (f(1.0) - p1))


# --------------------------------------------------------------------------- #
# input-side image preparation (utils/read_tf_records.py:88-112)
# --------------------------------------------------------------------------- #
_BICUBIC_TABLE = 1 << 10


def _bicubic_coeffs():
    """resize_bicubic_op.cc InitCoeffsTable [TF-upstream, recalled]: Keys kernel, A = -0.75, 1025 x 2 float32 entries."""
    A = -0.75
    tab = np.empty((_BICUBIC_TABLE + 1) * 2, np.float32)
    for i in range(_BICUBIC_TABLE + 1):
        x = float(np.float32(i * 1.0 / _BICUBIC_TABLE))
        tab[i * 2] = np.float32(((A + 2) * x - (A + 3)) * x * x + 1)
        x = float(np.float32(x + 1.0))
        tab[i * 2 + 1] = np.float32(((A * x - 5 * A) * x + 8 * A) * x - 4 * A)
    return tab


def _bicubic_taps(scale, out_size, limit, tab):
    """GetWeightsAndIndices for every output location: weights [out,4] float32, indices [out,4]."""
    loc = np.arange(out_size, dtype=np.float32)
    pos = (np.float32(scale) * loc).astype(np.float32)
    in_loc = pos.astype(np.int64)                                      # int64 in_loc = scale * out_loc
    delta = (pos - in_loc.astype(np.float32)).astype(np.float32)
    offset = np.rint((delta * np.float32(_BICUBIC_TABLE)).astype(np.float32)).astype(np.int64)      # lrintf: round half to even
    w = np.stack([tab[offset * 2 + 1], tab[offset * 2], tab[(_BICUBIC_TABLE - offset) * 2], tab[(_BICUBIC_TABLE - offset) * 2 + 1]], 1)
    idx = np.clip(in_loc[:, None] + np.arange(-1, 3)[None, :], 0, limit - 1)
    return w.astype(np.float32), idx


def resize_bicubic_tf13(images, out_h, out_w):
    """tf.image.resize_bicubic(images [B,H,W,C], [out_h, out_w]) with align_corners=False as TensorFlow 1.3 computes
    it [TF-upstream, recalled]: legacy coordinates in = out * in_size / out_size, table-sampled Keys weights, clamped
    taps, horizontal pass then vertical pass, float32 sums in index order.  Returns float32."""
    x = np.asarray(images)
    B, H, W, C = x.shape
    tab = _bicubic_coeffs()
    wy, iy = _bicubic_taps(np.float32(H) / np.float32(out_h), out_h, H, tab)
    wx, ix = _bicubic_taps(np.float32(W) / np.float32(out_w), out_w, W, tab)
    xf = x.astype(np.float32)
    f = np.float32
    rows = xf[:, iy]                                                   # [B, out_h, 4, W, C]
    g = rows[:, :, :, ix]                                              # [B, out_h, 4, out_w, 4, C]
    wxe = wx[None, None, None, :, :, None]
    h = (g[..., 0, :] * wxe[..., 0, :]).astype(f)
    for k in range(1, 4):
        h = (h + (g[..., k, :] * wxe[..., k, :]).astype(f)).astype(f)  # [B, out_h, 4, out_w, C]
    wye = wy[None, :, :, None, None]
    out = (h[:, :, 0] * wye[:, :, 0]).astype(f)
    for k in range(1, 4):
        out = (out + (h[:, :, k] * wye[:, :, k]).astype(f)).astype(f)
    return out


def process_image(raw_u8, out_size):
    """read_tf_records.py:100-111 on a batch of decoded uint8 images [B,H0,W0,C]: central crop to min(H0,W0)
    (resize_image_with_crop_or_pad), resize_bicubic to [out_size, out_size], float32 / 255."""
    x = np.asarray(raw_u8)
    H0, W0 = x.shape[1], x.shape[2]
    crop = min(H0, W0)
    oy, ox = (H0 - crop) // 2, (W0 - crop) // 2
    x = x[:, oy:oy + crop, ox:ox + crop]
    return (resize_bicubic_tf13(x, out_size, out_size) / np.float32(255.0)).astype(np.float32)
