"""CPU tests that pin the oracle (oracle/): hand-derived known answers, float64 finite
differences, torch CPU cross-checks, the reference's rectangle fixture, and the committed
golden vectors.  SURVEY.md 8(c)."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import graph as G
from oracle import tf_ops as T

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden as mg  # noqa: E402


# ------------------------------------------------------------------ sampler known answers
KAT_IMG = np.array([[1, 2, 3], [4, 5, 6]], np.float32).reshape(1, 2, 3, 1)
KAT = [((0, 0), 1), ((2, 1), 6), ((0.5, 0.5), 3), ((1.25, 0.75), 4.5), ((-0.5, 0), 0.5), ((-1, 0), 0), ((2.5, 1), 3),
       ((3, 1), 0), ((2, 1.5), 3), ((2, 2), 0)]


@pytest.mark.parametrize("pt,expect", KAT)
def test_resampler_known_answers(pt, expect):
    w = np.array(pt, np.float32).reshape(1, 1, 2)
    assert T.resampler(KAT_IMG, w)[0, 0, 0] == pytest.approx(expect, abs=1e-6)


def test_resampler_near_minus_one():
    w = np.array([-0.999, -0.999], np.float32).reshape(1, 1, 2)
    assert T.resampler(KAT_IMG, w)[0, 0, 0] == pytest.approx(1e-6, rel=0.05)


def test_resampler_nan_is_zero():
    w = np.array([np.nan, 0.5], np.float32).reshape(1, 1, 2)
    assert T.resampler(KAT_IMG, w)[0, 0, 0] == 0.0
    gd, gw = T.resampler_grad(KAT_IMG, w, np.ones((1, 1, 1), np.float32))
    assert not gd.any() and not gw.any()


def test_quirk_zero_flow_transposes():
    """coords() stacks (Y,X) but the sampler reads channel 0 as x (tf_utils.py:48-51)."""
    ramp = np.arange(16, dtype=np.float32).reshape(1, 4, 4, 1)
    out = T.resample_layer(ramp, T.warp_pts_layer(np.zeros((1, 4, 4, 2), np.float32)))
    assert np.array_equal(out[0, :, :, 0], ramp[0, :, :, 0].T)
    c = T.coords(3, 5, 2)
    assert c.shape == (2, 3, 5, 2) and c[1, 2, 4, 0] == 2 and c[1, 2, 4, 1] == 4


def _grid_sample(data, warp):
    B, H, W, C = data.shape
    g = torch.as_tensor(warp).double().clone()
    g[..., 0] = 2 * g[..., 0] / (W - 1) - 1
    g[..., 1] = 2 * g[..., 1] / (H - 1) - 1
    d = torch.as_tensor(data).double().permute(0, 3, 1, 2)
    return d, g


def test_resampler_matches_grid_sample_float64():
    rng = np.random.default_rng(0)
    data = rng.random((2, 7, 9, 3))
    warp = rng.uniform(-2, [10, 8], size=(2, 5, 6, 2))
    go = rng.standard_normal((2, 5, 6, 3))
    out = T.resampler(data, warp)
    gd, gw = T.resampler_grad(data, warp, go)
    d, g = _grid_sample(data, warp)
    d.requires_grad_(True)
    wt = torch.as_tensor(warp).double().requires_grad_(True)
    gg = torch.stack([2 * wt[..., 0] / 8 - 1, 2 * wt[..., 1] / 6 - 1], -1)
    o = torch.nn.functional.grid_sample(d, gg, mode="bilinear", padding_mode="zeros", align_corners=True)
    o.backward(torch.as_tensor(go).permute(0, 3, 1, 2))
    assert np.allclose(out, o.detach().permute(0, 2, 3, 1).numpy(), atol=1e-12)
    assert np.allclose(gd, d.grad.permute(0, 2, 3, 1).numpy(), atol=1e-12)
    assert np.allclose(gw, wt.grad.numpy(), atol=1e-10)


def test_resampler_grad_finite_differences():
    rng = np.random.default_rng(1)
    data = rng.random((1, 6, 5, 2))
    warp = rng.uniform(0.2, [3.8, 4.8], size=(1, 4, 3, 2))
    warp = np.floor(warp) + 0.25 + 0.5 * rng.random(warp.shape)       # away from integer coordinates
    go = rng.standard_normal((1, 4, 3, 2))
    gd, gw = T.resampler_grad(data, warp, go)
    eps = 1e-6
    num = np.zeros_like(warp)
    for idx in np.ndindex(warp.shape):
        wp, wm = warp.copy(), warp.copy()
        wp[idx] += eps
        wm[idx] -= eps
        num[idx] = ((T.resampler(data, wp) - T.resampler(data, wm)) * go).sum() / (2 * eps)
    assert np.allclose(gw, num, rtol=1e-6, atol=1e-8)
    numd = np.zeros_like(data)
    for idx in np.ndindex(data.shape):
        dp = data.copy()
        dp[idx] += 1.0
        numd[idx] = ((T.resampler(dp, warp) - T.resampler(data, warp)) * go).sum()      # linear in data
    assert np.allclose(gd, numd, atol=1e-10)


def test_resampler_boundary_grad_rule():
    """x exactly -1 fails the strict test: zero output and zero gradient (differs from grid_sample)."""
    w = np.array([[-1.0, 0.5], [0.5, -1.0], [3.0, 0.5], [0.5, 2.0]], np.float32).reshape(1, 4, 2)
    out = T.resampler(KAT_IMG, w)
    gd, gw = T.resampler_grad(KAT_IMG, w, np.ones((1, 4, 1), np.float32))
    assert not out.any() and not gw.any() and not gd.any()


def test_rectangle_fixture(golden_dir):
    """The reference's test_resampler.py restated with assertions: the rectangle rotates by 10 degrees."""
    g = np.load(os.path.join(golden_dir, "rectangle.npz"))
    H, W = int(g["shape"][0]), int(g["shape"][1])
    out8, angle = mg.rectangle_run(H, W, tuple(int(v) for v in g["bounds"]), g["fg"], g["bg"])
    assert abs(angle - 10.0) < 0.5
    assert angle == pytest.approx(float(g["angle_deg"]), abs=1e-6)
    assert hashlib.sha256(out8.tobytes()).digest() == g["sha256"].tobytes()
    # upper clip equals W/H, i.e. lands out of range -> zeros (test_resampler.py:38-39)
    warp = mg.rectangle_warp(H, W)
    assert (out8[(warp[0, ..., 0] >= W) | (warp[0, ..., 1] >= H)] == 0).all()


def test_sampler_golden_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "sampler_kat.npz"))
    out = T.resampler(g["kat_img"], g["kat_pts"].reshape(1, -1, 2))[0, :, 0]
    exp = g["kat_expect"]
    ok = ~np.isnan(exp)
    assert np.allclose(out[ok], exp[ok], atol=1e-6)
    for C in (1, 3, 4):
        data, warp, go = g["c%d_data" % C], g["c%d_warp" % C], g["c%d_go" % C]
        assert np.array_equal(T.resampler(data, warp), g["c%d_out" % C])
        fx, fy, cx, cy, mask = T.resampler_indices(data.shape, warp)
        assert np.array_equal(np.stack([fx, fy, cx, cy], -1), g["c%d_idx" % C]) and np.array_equal(mask, g["c%d_mask" % C])
        gd, gw = T.resampler_grad(data, warp, go)
        assert np.array_equal(gw, g["c%d_gw" % C]) and np.allclose(gd, g["c%d_gd" % C], rtol=1e-6, atol=1e-7)


# ------------------------------------------------------------------ padding / conv / deconv
def test_same_pad_table():
    """on even sizes k5s2 -> (1,2), k3s2 -> (0,1), k5s1 -> (2,2), k3s1 -> (1,1)."""
    assert T.same_pad(224, 5, 2) == (112, 1, 2)
    assert T.same_pad(28, 3, 2) == (14, 0, 1)
    assert T.same_pad(112, 5, 1) == (112, 2, 2)
    assert T.same_pad(14, 3, 1) == (14, 1, 1)
    assert T.same_pad(7, 3, 2) == (4, 1, 1)


@pytest.mark.parametrize("k,s,H", [(5, 2, 12), (5, 1, 9), (3, 2, 14), (3, 1, 7), (3, 2, 7)])
def test_conv_matches_torch_explicit_padding(k, s, H):
    rng = np.random.default_rng(k * 10 + s)
    x = rng.standard_normal((2, H, H, 3))
    w = rng.standard_normal((k, k, 3, 4))
    b = rng.standard_normal(4)
    y = T.conv2d_same(x, w, b, s, s)
    _, pt, pb = T.same_pad(H, k, s)
    xt = torch.as_tensor(x).permute(0, 3, 1, 2).requires_grad_(True)
    wt = torch.as_tensor(w).requires_grad_(True)
    yt = torch.nn.functional.conv2d(torch.nn.functional.pad(xt, (pt, pb, pt, pb)), wt.permute(3, 2, 0, 1),
                                    torch.as_tensor(b), stride=s)
    assert np.allclose(y, yt.detach().permute(0, 2, 3, 1).numpy(), atol=1e-12)
    gy = rng.standard_normal(y.shape)
    yt.backward(torch.as_tensor(gy).permute(0, 3, 1, 2))
    gx, gw, gb = T.conv2d_same_grads(x, w, gy, s, s)
    assert np.allclose(gx, xt.grad.permute(0, 2, 3, 1).numpy(), atol=1e-12)
    assert np.allclose(gw, wt.grad.numpy(), atol=1e-10)
    assert np.allclose(gb, gy.sum((0, 1, 2)))


def test_symmetric_torch_padding_differs():
    """torch's padding=k//2 is NOT TF-SAME for stride 2 on even sizes."""
    rng = np.random.default_rng(3)
    x, w = rng.standard_normal((1, 8, 8, 2)), rng.standard_normal((5, 5, 2, 2))
    y = T.conv2d_same(x, w, None, 2, 2)
    yt = torch.nn.functional.conv2d(torch.as_tensor(x).permute(0, 3, 1, 2), torch.as_tensor(w).permute(3, 2, 0, 1), stride=2,
                                    padding=2)
    assert np.abs(y - yt.permute(0, 2, 3, 1).numpy()).max() > 0.1


@pytest.mark.parametrize("k,s,H", [(5, 2, 12), (3, 2, 14), (3, 1, 6), (5, 2, 8)])
def test_deconv_is_input_gradient_of_same_conv(k, s, H):
    rng = np.random.default_rng(k + s + H)
    h = -(-H // s)
    x = rng.standard_normal((2, h, h, 3))
    w = rng.standard_normal((k, k, 4, 3))
    y = T.conv2d_transpose_same(x, w, (2, H, H, 4), s, s)
    big = torch.zeros(2, 4, H, H, dtype=torch.float64, requires_grad=True)
    _, pt, pb = T.same_pad(H, k, s)
    out = torch.nn.functional.conv2d(torch.nn.functional.pad(big, (pt, pb, pt, pb)), torch.as_tensor(w).permute(3, 2, 0, 1),
                                     stride=s)
    out.backward(torch.as_tensor(x).permute(0, 3, 1, 2))
    assert np.allclose(y, big.grad.permute(0, 2, 3, 1).numpy(), atol=1e-12)
    # crop form used by the torch-CPU port
    yt = G.TorchCpuOps(torch.float64).deconv(torch.as_tensor(x).permute(0, 3, 1, 2), torch.as_tensor(w), (2, H, H, 4), s)
    assert np.allclose(y, yt.permute(0, 2, 3, 1).numpy(), atol=1e-12)
    # gradients against autograd of the crop form
    xt = torch.as_tensor(x).permute(0, 3, 1, 2).requires_grad_(True)
    wt = torch.as_tensor(w).requires_grad_(True)
    gy = rng.standard_normal(y.shape)
    G.TorchCpuOps(torch.float64).deconv(xt, wt, (2, H, H, 4), s).backward(torch.as_tensor(gy).permute(0, 3, 1, 2))
    gx, gw = T.conv2d_transpose_same_grads(x, w, gy, s, s)
    assert np.allclose(gx, xt.grad.permute(0, 2, 3, 1).numpy(), atol=1e-10)
    assert np.allclose(gw, wt.grad.numpy(), atol=1e-10)


def test_layer_golden_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "layers.npz"))
    for name, k, s in [("k5s2", 5, 2), ("k5s1", 5, 1), ("k3s2", 3, 2), ("k3s1", 3, 1)]:
        y = T.conv2d_same(g["conv_%s_x" % name], g["conv_%s_w" % name], g["conv_%s_b" % name], s, s)
        assert np.allclose(y, g["conv_%s_y" % name], rtol=1e-6, atol=1e-6)
        H = g["deconv_%s_y" % name].shape[1]
        yd = T.conv2d_transpose_same(g["deconv_%s_x" % name], g["deconv_%s_w" % name], g["deconv_%s_y" % name].shape, s, s)
        assert np.allclose(yd, g["deconv_%s_y" % name], rtol=1e-6, atol=1e-6) and yd.shape[1] == H
    assert np.allclose(T.linear(g["lin_x"], g["lin_m"], g["lin_b"]), g["lin_y"], rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------ activations / losses / Adam
def test_activations_and_losses():
    x = np.array([-2.0, -0.5, 0.0, 0.5, 3.0], np.float32)
    assert np.allclose(T.lrelu(x), np.where(x > 0, x, 0.2 * x))
    assert np.allclose(T.relu(x), np.maximum(x, 0))
    assert np.allclose(T.lrelu_grad(x, np.ones_like(x)), [0.2, 0.2, 0.6, 1, 1])
    a = np.arange(24, dtype=np.float32).reshape(1, 2, 3, 4) / 10
    b = a[..., ::-1].copy()
    assert T.euclidean_loss(a, b) == pytest.approx(((a - b) ** 2).sum() / 6)
    assert T.l1_loss(a, b) == pytest.approx(np.abs(a - b).sum() / 6)
    eps = 1e-3
    ga = T.euclidean_loss_grad(a, b)
    ap = a.copy(); ap[0, 1, 2, 3] += eps
    assert (T.euclidean_loss(ap, b) - T.euclidean_loss(a, b)) / eps == pytest.approx(ga[0, 1, 2, 3], rel=1e-2)


def test_adam_is_tf_flavoured():
    rng = np.random.default_rng(5)
    th = rng.standard_normal(100).astype(np.float32)
    m = np.zeros(100, np.float32); v = np.zeros(100, np.float32)
    th_t = torch.tensor(th.copy(), requires_grad=True)
    opt = torch.optim.Adam([th_t], lr=1e-3, eps=1e-8)
    th0 = th.copy()
    for t in range(1, 4):
        g = rng.standard_normal(100).astype(np.float32) * 1e-4     # small grads: eps placement matters
        th, m, v = T.adam_tf_step(th, g, m, v, t, 1e-3)
        th_t.grad = torch.tensor(g)
        opt.step()
    # closed form for step 1 on a fresh state
    g1 = np.float32(0.01)
    t1, _, _ = T.adam_tf_step(np.float32([1.0]), np.float32([g1]), np.zeros(1, np.float32), np.zeros(1, np.float32), 1, 1e-3)
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert t1[0] == pytest.approx(1.0 - lr_t * (0.1 * g1) / (np.sqrt(0.001 * g1 * g1) + 1e-8), rel=1e-6)
    assert np.abs(th - th0).max() > 0
    # eps outside the bias correction: close to, but not the same as, torch.optim.Adam
    diff = np.abs(th - th_t.detach().numpy()).max()
    assert 0 < diff < 2e-3
    assert float(T.adam_lr_t(1e-4, 1)) == pytest.approx(1e-4 * np.sqrt(0.001) / 0.1, rel=1e-5)


# ------------------------------------------------------------------ graphs
@pytest.mark.parametrize("kind", ["base", "highdim", "lowdim", "tinghui"])
def test_graph_numpy_vs_torch_cpu(kind):
    H, V, B = 32, 19, 2
    P = G.init_params(G.appflow_param_shapes(H, V, kind), 0)
    rng = np.random.default_rng(1)
    img = rng.random((B, H, H, 3), dtype=np.float32)
    disp = np.eye(V, dtype=np.float32)[[3, 7]]
    o = G.appearance_flow_forward(G.NumpyOps(), P, img, disp, kind)
    t = G.appearance_flow_forward(G.TorchCpuOps(), P, img, disp, kind)
    assert np.allclose(o["flow_field"], t["flow_field"].numpy(), atol=1e-5)
    assert np.allclose(o["gen"], t["gen"].numpy(), atol=1e-4)


def test_param_count_matches_survey():
    n = sum(int(np.prod(s[1])) for s in G.appflow_param_shapes(224, 2, "base").values())
    assert n == 138749696            # 138.75 M at 224^2
    n128 = sum(int(np.prod(s[1])) for s in G.appflow_param_shapes(128, 2, "base").values())
    assert abs(n128 - 69.5e6) < 0.1e6


def test_colordepth_graph_runs_and_split_order():
    conf = {"use_color": "", "use_depth": "", "depth_lr_factor": 0.1}
    H, B, V = 32, 1, 2
    P = G.init_params(G.colordepth_param_shapes(H, V, conf), 0)
    rng = np.random.default_rng(2)
    img, dimg = rng.random((B, H, H, 3), dtype=np.float32), rng.random((B, H, H, 1), dtype=np.float32)
    disp = rng.standard_normal((B, V)).astype(np.float32)
    o = G.colordepth_forward(G.NumpyOps(), P, conf, img, dimg, disp)
    t = G.colordepth_forward(G.TorchCpuOps(), P, conf, img, dimg, disp)
    assert o["gen_image1"].shape == (B, H, H, 3) and o["gen_dimage1"].shape == (B, H, H, 1)
    assert np.allclose(o["gen_image1"], t["gen_image1"].numpy(), atol=1e-4)
    l_np = G.colordepth_loss(G.NumpyOps(), o, conf, img, dimg)
    l_t = G.colordepth_loss(G.TorchCpuOps(), t, conf, img, dimg)
    assert float(l_np) == pytest.approx(float(l_t), rel=1e-4)


def test_fusion_definition():
    rng = np.random.default_rng(3)
    gens = rng.random((4, 1, 3, 3, 3))
    logits = rng.standard_normal((4, 1, 3, 3, 1))
    f = G.fuse_views(G.NumpyOps(), gens, logits)
    w = np.exp(logits) / np.exp(logits).sum(0, keepdims=True)
    assert np.allclose(f, (w * gens).sum(0))


MO_CONFS = [{"use_color": "", "use_depth": 0.1, "combination_image": "", "gen_sep_images": "", "predict_target_masks": 0.1,
             "masked_image_loss": ""},
            {"use_color": "", "combination_image": "", "fully_conv": ""}]


def _mo_batch(rng, B, H):
    names = ["image0", "image0_mask0", "image0_mask1", "image1", "image1_only0", "image1_only1", "image1_mask0", "image1_mask1",
             "depth0", "depth1", "depth1_only0", "depth1_only1"]
    b = {n: rng.random((B, H, H, 3 if (n.startswith("image") and "mask" not in n) else 1), dtype=np.float32) for n in names}
    b["displacement"] = rng.standard_normal((B, 2)).astype(np.float32)
    return b


@pytest.mark.parametrize("conf", MO_CONFS)
def test_multiobject_graph_numpy_vs_torch_cpu(conf):
    """The two backends of the oracle agree on the restatement of multiobject_appflow.py:123-283, the heads come out in
    the reference's pop() order, and a masked loss ignores pixels outside the mask."""
    rng = np.random.default_rng(3)
    B, H, V = 1, 32, 2
    P = G.init_params(G.multiobject_param_shapes(H, V, conf), 0)
    b = _mo_batch(rng, B, H)
    o = G.multiobject_forward(G.NumpyOps(), P, conf, b)
    t = G.multiobject_forward(G.TorchCpuOps(), P, conf, b)
    assert list(o) == [h[0] for h in G._mo_heads(conf)]
    for k in o:
        assert o[k].shape == (B, H, H, 3 if k.startswith("gen_image1") and "mask" not in k else 1)
        assert np.allclose(o[k], np.asarray(t[k]), rtol=1e-3, atol=1e-4), k
    l_np = float(G.multiobject_loss(G.NumpyOps(), o, conf, b))
    l_t = float(G.multiobject_loss(G.TorchCpuOps(), t, conf, b))
    assert l_np == pytest.approx(l_t, rel=1e-3)
    if "masked_image_loss" in conf:
        b2 = dict(b)
        b2["image1_only0"] = b["image1_only0"] + 5.0 * (1.0 - np.round(b["image1_mask0"]))     # change pixels where mask < 0.5 only
        b2["image1_mask0"] = np.round(b["image1_mask0"])
        b1 = dict(b, image1_mask0=np.round(b["image1_mask0"]))
        assert float(G.multiobject_loss(G.NumpyOps(), o, conf, b1)) == pytest.approx(float(G.multiobject_loss(G.NumpyOps(), o, conf, b2)), rel=1e-12)


def _mv_batch(rng, Vw, B, H, V, depth=True):
    b = {"image0": rng.random((Vw, B, H, H, 3), dtype=np.float32), "image0_mask0": rng.random((Vw, B, H, H, 1), dtype=np.float32),
         "image0_mask1": rng.random((Vw, B, H, H, 1), dtype=np.float32), "displacement": rng.standard_normal((Vw, B, V)).astype(np.float32),
         "image1": rng.random((B, H, H, 3), dtype=np.float32)}
    if depth:
        b["depth0"] = rng.random((Vw, B, H, H, 1), dtype=np.float32)
    return b


def test_multiview_graph_numpy_vs_torch_cpu_and_single_view_limit():
    """Config 5's definition (SURVEY 8(f)-3 on the multi-object trunk): backends agree; the fusion is a convex
    combination; with ONE view it is the identity (softmax of one logit = 1), so the model reduces to a multi-object
    flow decoder whose head ignores its third channel."""
    rng = np.random.default_rng(4)
    B, H, V, Vw = 1, 32, 2, 3
    conf = {"use_depth": 0.1}
    P = G.init_params(G.multiview_param_shapes(H, V, conf), 0)
    assert P["dec_image1/d0/w"].shape == (5, 5, 3, 32) and "pre_mask0_ob1/e0/w" in P and "pre_dimage0_f/e0/w" in P
    b = _mv_batch(rng, Vw, B, H, V)
    o = G.multiview_forward(G.NumpyOps(), P, conf, b)
    t = G.multiview_forward(G.TorchCpuOps(), P, conf, b)
    assert np.allclose(o["fused"], np.asarray(t["fused"]), rtol=1e-3, atol=1e-4)
    assert np.allclose(o["logits"], np.asarray(t["logits"]), rtol=1e-3, atol=1e-4)
    w = np.exp(o["logits"] - o["logits"].max(0, keepdims=True))
    w /= w.sum(0, keepdims=True)
    assert np.allclose(o["fused"], (w[..., None] * o["gens"]).sum(0), atol=1e-6) and np.allclose(w.sum(0), 1.0)
    one = G.multiview_forward(G.NumpyOps(), P, conf, {k: (v[:1] if k != "image1" else v) for k, v in b.items()})
    assert np.array_equal(one["fused"], one["gens"][0])
    # the same frame through the multi-object graph with a 2-channel flow head made of the head's first two channels
    mo_conf = {"use_color": "", "use_depth": 0.1, "combination_image": ""}
    P2 = {k: v for k, v in P.items()}
    P2["dec_image1/d0/w"] = P["dec_image1/d0/w"][:, :, :2, :]
    P2["d3_0/w"] = np.concatenate([P["d3_0/w"], P["d3_0/w"]], axis=3)          # two heads (colour flow, depth tanh): the colour
    P2["d3_0/b"] = np.concatenate([P["d3_0/b"], P["d3_0/b"]])                  # head pops the LAST group = a copy of ours
    for k in list(P):
        if k.startswith("dec_image1/"):
            P2["dec_dimage1_f/" + k.split("/", 1)[1]] = P[k][:, :, :1, :] if k.endswith("d0/w") else P[k]
    mo = {"image0": b["image0"][0], "depth0": b["depth0"][0], "image0_mask0": b["image0_mask0"][0], "image0_mask1": b["image0_mask1"][0],
          "displacement": b["displacement"][0]}
    single = G.multiobject_forward(G.NumpyOps(), P2, mo_conf, mo)
    assert np.allclose(one["gens"][0], single["gen_image1"], atol=1e-6)


def test_bf16_aware_backend_rounds_where_the_cuda_path_stores_bf16():
    """Bf16TorchCpuOps: layer outputs are bf16-representable, fp32 heads are not rounded, weight gradients stay fp32,
    and on bf16-representable data with a single layer it equals the fp32 backend up to the output rounding."""
    import torch
    rng = np.random.default_rng(9)
    ops16, ops32 = G.Bf16TorchCpuOps(), G.TorchCpuOps()
    r = lambda a: torch.as_tensor(a).to(torch.bfloat16).to(torch.float32)
    x = r(rng.standard_normal((2, 8, 8, 4)).astype(np.float32))
    w = r(rng.standard_normal((3, 3, 4, 8)).astype(np.float32)).requires_grad_(True)
    bias = torch.as_tensor(rng.standard_normal(8).astype(np.float32))
    y16 = ops16.lrelu(ops16.conv(ops16.from_nhwc(x), w, bias, 1))
    y32 = ops32.lrelu(ops32.conv(ops32.from_nhwc(x), w, bias, 1))
    assert torch.equal(y16, r(y16)) and torch.equal(y16, r(y32))
    head = ops16.conv(ops16.from_nhwc(x), w, bias, 1)                   # no activation: an fp32 head
    assert torch.equal(head, ops32.conv(ops32.from_nhwc(x), w, bias, 1)) and not torch.equal(head, r(head))
    g = torch.as_tensor(rng.standard_normal(tuple(y16.shape)).astype(np.float32))
    (gw,) = torch.autograd.grad(y16, w, g)
    assert not torch.equal(gw, r(gw))                                   # fp32 weight gradient
    # the gradient that reaches the conv is dPre = round(g * act'(y)) (the factor is applied before the one rounding:
    # the consumer's dgrad epilogue does it on the fp32 accumulator): compare with that fed to the fp32 backend
    y32b = ops32.conv(ops32.from_nhwc(x), w, bias, 1)
    slope = torch.where(y32b > 0, torch.ones_like(y32b), torch.full_like(y32b, 0.2))
    (gw_ref,) = torch.autograd.grad(y32b, w, r(g * slope))
    assert torch.allclose(gw, gw_ref, rtol=1e-5, atol=1e-6)


def test_graph_golden_vectors(golden_dir):
    """oracle/graph.py against the committed outputs of tests/golden/make_golden.py::graph_cases (whole graphs at 32x32:
    single-view app-flow, colour+depth, multi-object with and without the fully-conv bottleneck, 3-view fusion)."""
    import make_golden as M
    g = np.load(os.path.join(golden_dir, "graphs.npz"))
    out = M.graph_outputs()
    assert sorted(out) == sorted(g.files)
    for k, v in out.items():
        if k.endswith("_loss"):
            assert float(v) == pytest.approx(float(g[k]), rel=1e-6), k
        else:
            assert np.allclose(v, g[k].astype(np.float32), rtol=2e-3, atol=2e-3), k      # stored as float16
