"""GPU parity tests of conv / deconv / linear (forward, dgrad, wgrad), activations, the fused
loss (+ multi-view fusion) and Adam through the C ABI against the oracle.  Operands are
bf16-representable so the only differences are fp32 summation order and the bf16 rounding
of outputs: forward/activation tolerance 1e-2 relative (BASELINE), fp32 outputs 1e-4."""
import os

import numpy as np
import pytest
import torch

from oracle import tf_ops as T

pytestmark = pytest.mark.gpu

ALGOS = ["simt", "auto"]


def _store():
    from dynamic_multiview_3d_b200.variables import VariableStore
    return VariableStore(torch.device("cuda:0"))


def _var(store, name, arr):
    v = store.get(name, arr.shape, "zeros")
    v.master.copy_(torch.from_numpy(np.ascontiguousarray(arr)).cuda())
    store._cast(v.master, v.half, v.numel)
    return v


def _t(a, dtype=torch.bfloat16):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda().to(dtype)


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def bf16_round(x):
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name,k,s", [("k5s2", 5, 2), ("k5s1", 5, 1), ("k3s2", 3, 2), ("k3s1", 3, 1)])
def test_conv_golden(golden_dir, name, k, s, algo):
    from dynamic_multiview_3d_b200 import functional as F
    g = np.load(os.path.join(golden_dir, "layers.npz"))
    x, w, b, y, gy = (g["conv_%s_%s" % (name, t)] for t in ("x", "w", "b", "y", "gy"))
    st = _store()
    wv, bv = _var(st, "w", w), _var(st, "b", b)
    for xin in ([_t(x), _t(x, torch.float32)] if x.shape[-1] == 3 else [_t(x)]):
        xt = xin.requires_grad_(True)
        yf = F.conv2d(xt, wv, bv, s, None, algo, torch.float32)
        assert _rel(yf.detach().cpu().numpy(), y) < 1e-4
        yb = F.conv2d(xt, wv, bv, s, None, algo)
        assert _rel(yb.detach().float().cpu().numpy(), y) < 1e-2
        yb.backward(_t(gy))
        assert _rel(xt.grad.float().cpu().numpy(), g["conv_%s_gx" % name]) < 1e-2
        assert _rel(wv.grad.cpu().numpy(), g["conv_%s_gw" % name]) < 1e-4
        assert _rel(bv.grad.cpu().numpy(), g["conv_%s_gb" % name]) < 1e-4
    # fused activation == oracle lrelu on the fp32 pre-activation, then its gradient
    xt = _t(x).requires_grad_(True)
    ya = F.conv2d(xt, wv, bv, s, "lrelu", algo)
    assert _rel(ya.detach().float().cpu().numpy(), T.lrelu(y)) < 1e-2
    ya.backward(_t(gy))
    ya_np = ya.detach().float().cpu().numpy()
    gpre = bf16_round(gy * np.where(ya_np > 0, 1.0, np.where(ya_np < 0, 0.2, 0.6)))
    gx, gw, gb = T.conv2d_same_grads(x, w, gpre, s, s)
    assert _rel(wv.grad.cpu().numpy(), gw) < 1e-3 and _rel(bv.grad.cpu().numpy(), gb) < 1e-3
    assert _rel(xt.grad.float().cpu().numpy(), gx) < 1e-2


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name,k,s", [("k5s2", 5, 2), ("k5s1", 5, 1), ("k3s2", 3, 2), ("k3s1", 3, 1)])
def test_deconv_golden(golden_dir, name, k, s, algo):
    from dynamic_multiview_3d_b200 import functional as F
    g = np.load(os.path.join(golden_dir, "layers.npz"))
    x, w, y, gy = (g["deconv_%s_%s" % (name, t)] for t in ("x", "w", "y", "gy"))
    st = _store()
    wv = _var(st, "w", w)
    for odt, tol in [(torch.float32, 1e-4), (torch.bfloat16, 1e-2)]:
        xt = _t(x).requires_grad_(True)
        yo = F.deconv2d(xt, wv, y.shape[1:3], s, None, algo, odt)
        assert _rel(yo.detach().float().cpu().numpy(), y) < tol
        yo.backward(_t(gy, odt))
        assert _rel(xt.grad.float().cpu().numpy(), g["deconv_%s_gx" % name]) < 1e-2
        assert _rel(wv.grad.cpu().numpy(), g["deconv_%s_gw" % name]) < 1e-4


@pytest.mark.parametrize("algo", ALGOS)
def test_linear_golden(golden_dir, algo):
    from dynamic_multiview_3d_b200 import functional as F
    g = np.load(os.path.join(golden_dir, "layers.npz"))
    st = _store()
    mv, bv = _var(st, "Matrix", g["lin_m"]), _var(st, "b", g["lin_b"])
    xt = _t(g["lin_x"]).requires_grad_(True)
    y = F.linear(xt, mv, bv, None, algo)
    assert _rel(y.detach().float().cpu().numpy(), g["lin_y"]) < 1e-2
    y.backward(_t(g["lin_gy"]))
    assert _rel(xt.grad.float().cpu().numpy(), g["lin_gx"]) < 1e-2
    assert _rel(mv.grad.cpu().numpy(), g["lin_gm"]) < 1e-4
    assert _rel(bv.grad.cpu().numpy(), g["lin_gb"]) < 1e-4


REF_LAYERS = [  # (kind, k, s, H_in_of_big_side, cin, cout) of appearance_flow_model.py:88-125 at reduced spatial size
    ("conv", 5, 2, 32, 3, 32), ("conv", 5, 1, 16, 32, 32), ("conv", 5, 2, 16, 32, 64), ("conv", 5, 1, 8, 64, 64),
    ("conv", 3, 2, 8, 64, 128), ("conv", 3, 1, 4, 128, 128), ("conv", 3, 2, 14, 128, 256), ("conv", 3, 1, 7, 256, 256),
    ("conv", 5, 1, 16, 32, 64),
    ("deconv", 3, 2, 14, 256, 128), ("deconv", 3, 2, 8, 128, 64), ("deconv", 5, 2, 16, 64, 32), ("deconv", 5, 2, 32, 32, 2),
]


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("kind,k,s,H,cin,cout", REF_LAYERS)
def test_reference_layer_shapes(kind, k, s, H, cin, cout, algo):
    """Every (kernel, stride, channel) combination of the reference graph, seeded inputs."""
    from dynamic_multiview_3d_b200 import functional as F
    rng = np.random.default_rng(k * 1000 + s * 100 + cin + cout)
    B = 3
    st = _store()
    if kind == "conv":
        x = bf16_round(rng.standard_normal((B, H, H, cin)))
        w = bf16_round(rng.standard_normal((k, k, cin, cout)) * T.conv_stddev(k, k, cin))
        b = rng.standard_normal(cout).astype(np.float32)
        y = T.conv2d_same(x, w, b, s, s)
        gy = bf16_round(rng.standard_normal(y.shape))
        gx, gw, gb = T.conv2d_same_grads(x, w, gy, s, s)
        wv, bv = _var(st, "w", w), _var(st, "b", b)
        xt = _t(x).requires_grad_(True)
        yo = F.conv2d(xt, wv, bv, s, None, algo)
        assert _rel(yo.detach().float().cpu().numpy(), y) < 1e-2
        yo.backward(_t(gy))
        assert _rel(bv.grad.cpu().numpy(), gb) < 1e-4
    else:
        h = -(-H // s)
        x = bf16_round(rng.standard_normal((B, h, h, cin)))
        w = bf16_round(rng.standard_normal((k, k, cout, cin)) * T.deconv_stddev(k, k, cin, s, s))
        y = T.conv2d_transpose_same(x, w, (B, H, H, cout), s, s)
        gy = bf16_round(rng.standard_normal(y.shape))
        gx, gw = T.conv2d_transpose_same_grads(x, w, gy, s, s)
        wv = _var(st, "w", w)
        xt = _t(x).requires_grad_(True)
        yo = F.deconv2d(xt, wv, (H, H), s, None, algo)
        assert _rel(yo.detach().float().cpu().numpy(), y) < 1e-2
        yo.backward(_t(gy))
    assert _rel(xt.grad.float().cpu().numpy(), gx) < 1e-2
    assert _rel(wv.grad.cpu().numpy(), gw) < 2e-4


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("M,K,N", [(64, 12544, 256), (8, 4160, 4096), (64, 19, 64), (3, 64, 64), (64, 256, 12544)])
def test_linear_shapes(M, K, N, algo):
    from dynamic_multiview_3d_b200 import functional as F
    rng = np.random.default_rng(M + K + N)
    x = bf16_round(rng.standard_normal((M, K)))
    m = bf16_round(rng.standard_normal((K, N)) * T.linear_stddev(K))
    b = rng.standard_normal(N).astype(np.float32)
    y = T.linear(x, m, b)
    gy = bf16_round(rng.standard_normal(y.shape))
    gx, gm, gb = T.linear_grads(x, m, gy)
    st = _store()
    mv, bv = _var(st, "Matrix", m), _var(st, "b", b)
    xt = _t(x).requires_grad_(True)
    yo = F.linear(xt, mv, bv, "lrelu", algo)
    assert _rel(yo.detach().float().cpu().numpy(), T.lrelu(y)) < 1e-2
    yo2 = F.linear(xt, mv, bv, None, algo)
    yo2.backward(_t(gy))
    assert _rel(xt.grad.float().cpu().numpy(), gx) < 1e-2
    assert _rel(mv.grad.cpu().numpy(), gm) < 2e-4
    assert _rel(bv.grad.cpu().numpy(), gb) < 1e-4


@pytest.mark.parametrize("act", ["lrelu", "relu", "tanh"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_activations(act, dtype):
    from dynamic_multiview_3d_b200 import functional as F
    rng = np.random.default_rng(1)
    x = bf16_round(rng.standard_normal(1000 + 3))
    x[:3] = [0.0, -1.0, 2.0]
    ref = {"lrelu": T.lrelu, "relu": T.relu, "tanh": np.tanh}[act](x)
    xt = _t(x, dtype).requires_grad_(True)
    y = F.activation(xt, act)
    tol = 1e-6 if dtype == torch.float32 else 1e-2
    assert _rel(y.detach().float().cpu().numpy(), ref) < tol
    y.backward(torch.ones_like(y))
    slope = {"lrelu": np.where(x > 0, 1.0, np.where(x < 0, 0.2, 0.6)), "relu": (x > 0).astype(np.float32),
             "tanh": 1 - np.tanh(x) ** 2}[act]
    assert _rel(xt.grad.float().cpu().numpy(), slope) < (1e-5 if dtype == torch.float32 else 2e-2)


@pytest.mark.parametrize("mode", ["l2", "l1"])
@pytest.mark.parametrize("C,weights", [(3, None), (4, [1, 1, 1, 0.1]), (1, [0.1])])
def test_fused_loss(mode, C, weights):
    from dynamic_multiview_3d_b200 import functional as F
    rng = np.random.default_rng(C)
    a = rng.random((5, 33, 17, C), dtype=np.float32)
    b = rng.random((5, 33, 17, C), dtype=np.float32)
    a[0, 0, 0] = b[0, 0, 0]                                  # exact zeros: sign(0) = 0
    w = np.ones(C, np.float32) if weights is None else np.asarray(weights, np.float32)
    loss_fn, grad_fn = (T.euclidean_loss, T.euclidean_loss_grad) if mode == "l2" else (T.l1_loss, T.l1_loss_grad)
    ref = sum(w[c] * loss_fn(a[..., c:c + 1], b[..., c:c + 1]) for c in range(C))
    gref = np.concatenate([w[c] * grad_fn(a[..., c:c + 1], b[..., c:c + 1]) for c in range(C)], -1)
    at = torch.from_numpy(a).cuda().requires_grad_(True)
    loss = F.reconstruction_loss(at, torch.from_numpy(b).cuda(), mode, weights)
    assert float(loss.detach()) == pytest.approx(ref, rel=1e-5)
    (loss * 0.5).backward()                                  # upstream scalar chained on the device
    assert np.allclose(at.grad.cpu().numpy(), 0.5 * gref, rtol=1e-5, atol=1e-12)
    # deterministic
    l2 = F.reconstruction_loss(at.detach(), torch.from_numpy(b).cuda(), mode, weights)
    assert float(l2) == float(loss)


def test_masked_loss_and_view_fusion():
    from dynamic_multiview_3d_b200 import functional as F
    rng = np.random.default_rng(2)
    V, B, H, W, C = 4, 2, 19, 23, 3
    gens = rng.random((V, B, H, W, C), dtype=np.float32)
    logits = rng.standard_normal((V, B, H, W)).astype(np.float32)
    tgt = rng.random((B, H, W, C), dtype=np.float32)
    mask = (rng.random((B, H, W, 1)) > 0.4).astype(np.float32)
    gt = torch.from_numpy(gens).cuda().requires_grad_(True)
    lt = torch.from_numpy(logits).cuda().requires_grad_(True)
    loss, fused = F.fused_views_loss(gt, lt, torch.from_numpy(tgt).cuda(), "l2", mask=torch.from_numpy(mask).cuda())
    loss.backward()
    g64 = torch.from_numpy(gens).double().requires_grad_(True)
    l64 = torch.from_numpy(logits).double().requires_grad_(True)
    f64 = (torch.softmax(l64, 0)[..., None] * g64).sum(0)
    d = (f64 - torch.from_numpy(tgt).double()) * torch.from_numpy(mask).double()
    ref = (d * d).sum(-1).mean()
    ref.backward()
    assert np.allclose(fused.cpu().numpy(), f64.detach().numpy(), atol=1e-6)
    assert float(loss) == pytest.approx(float(ref), rel=1e-5)
    assert np.allclose(gt.grad.cpu().numpy(), g64.grad.numpy(), rtol=1e-4, atol=1e-9)
    assert np.allclose(lt.grad.cpu().numpy(), l64.grad.numpy(), rtol=1e-3, atol=1e-9)


def test_adam_matches_tf_oracle():
    from dynamic_multiview_3d_b200.optimizer import TFAdam
    rng = np.random.default_rng(3)
    st = _store()
    shapes = [(5, 5, 3, 32), (32,), (1000, 37), (37,), (3,)]
    vs = [st.get("v%d" % i, s, "normal", 0.1) for i, s in enumerate(shapes)]
    st.finalize()
    opt = TFAdam(st, 1e-3)
    ref = [(v.master.cpu().numpy().copy(), np.zeros(v.shape, np.float32), np.zeros(v.shape, np.float32)) for v in vs]
    for t in range(1, 4):
        gs = [rng.standard_normal(s).astype(np.float32) * 10.0 ** rng.integers(-6, 0) for s in shapes]
        for v, g in zip(vs, gs):
            v.grad.copy_(torch.from_numpy(g).cuda())
        opt.step()
        ref = [T.adam_tf_step(th, g, m, vv, t, 1e-3) for (th, m, vv), g in zip(ref, gs)]
    assert opt.t == 3
    for v, (th, m, vv) in zip(vs, ref):
        assert np.allclose(v.master.cpu().numpy(), th, rtol=1e-5, atol=1e-7)
        assert np.allclose(v.m.cpu().numpy(), m, rtol=1e-5, atol=1e-12)       # fp32 FMA contraction vs NumPy
        assert np.allclose(v.v.cpu().numpy(), vv, rtol=1e-5, atol=1e-15)
        assert torch.equal(v.half, v.master.to(torch.bfloat16))          # bf16 compute copy refreshed in the same pass


@pytest.mark.parametrize("M,K,N", [(64, 192, 256), (5, 72, 136), (130, 64, 128), (64, 4160, 512), (16, 128, 12544)])
def test_linear_wgrad_adam_fused(M, K, N):
    """dmv_linear_wgrad_adam: the weight gradient formed in registers equals x^T dy (oracle, float64 accumulate), and the
    update applied to theta / m / v / the bf16 copy is ApplyAdam on exactly that gradient -- two consecutive steps, so
    non-zero moments are read back.  Shapes cover row / column tails of the 64 x 128 tile and several sample chunks."""
    from dynamic_multiview_3d_b200 import _lib
    rng = np.random.default_rng(M * 7 + K + N)
    theta = (rng.standard_normal((K, N)) * 0.05).astype(np.float32)
    th, m, v = (torch.from_numpy(theta).cuda(), torch.zeros(K, N, device="cuda"), torch.zeros(K, N, device="cuda"))
    half = torch.zeros(K, N, dtype=torch.bfloat16, device="cuda")
    dw = torch.empty(K, N, device="cuda")
    state = torch.tensor([1.0, 1.0, 0.0, 0.0], device="cuda")
    ref = (theta.copy(), np.zeros((K, N), np.float32), np.zeros((K, N), np.float32))
    st = torch.cuda.current_stream().cuda_stream
    for t in (1, 2):
        x = bf16_round(rng.standard_normal((M, K)))
        dy = bf16_round(rng.standard_normal((M, N)) * 10.0 ** rng.integers(-4, 0))
        xt, dyt = _t(x), _t(dy)
        _lib.call("dmv_adam_tick", state.data_ptr(), 1e-3, 0.9, 0.999, st)
        _lib.call("dmv_linear_wgrad_adam", xt.data_ptr(), dyt.data_ptr(), th.data_ptr(), m.data_ptr(), v.data_ptr(), half.data_ptr(),
                  dw.data_ptr(), M, K, N, state.data_ptr(), 0.9, 0.999, 1e-8, 1.0, st)
        torch.cuda.synchronize()
        g = dw.cpu().numpy()
        gref = (x.astype(np.float64).T @ dy.astype(np.float64)).astype(np.float32)
        assert _rel(g, gref) < 1e-5
        ref = T.adam_tf_step(ref[0], g, ref[1], ref[2], t, 1e-3)
        # the kernel contracts m + (g - m)(1 - b1) into one FMA where NumPy rounds twice: a few ulp, amplified where g and m
        # cancel -- so the moments are held relative to their scale, theta to a thousandth of one step (lr = 1e-3)
        assert _rel(m.cpu().numpy(), ref[1]) < 1e-6
        assert _rel(v.cpu().numpy(), ref[2]) < 1e-6
        assert float(np.abs(th.cpu().numpy() - ref[0]).max()) < 1e-6
        assert torch.equal(half, th.to(torch.bfloat16))


def test_u8_to_f32_matches_reference_division():
    """dmv_u8_to_f32: float32(pixel) / 255 in IEEE division, bit for bit (read_tf_records.py:111)."""
    from dynamic_multiview_3d_b200 import _lib
    src = torch.arange(256, dtype=torch.uint8).repeat(64).cuda()
    dst = torch.empty(src.numel(), dtype=torch.float32, device="cuda")
    _lib.call("dmv_u8_to_f32", src.data_ptr(), dst.data_ptr(), src.numel(), 255.0, torch.cuda.current_stream().cuda_stream)
    ref = (np.arange(256, dtype=np.uint8).astype(np.float32) / np.float32(255.0))
    assert np.array_equal(dst.cpu().numpy().reshape(64, 256), np.tile(ref, (64, 1)))
