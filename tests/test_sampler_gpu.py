"""GPU parity tests of the bilinear sampler through the C ABI (via functional.py) against the
oracle: corner indices and validity masks bit-exact; outputs and grad_warp bit-equal to the
NumPy restatement (the kernel keeps its summation order with non-contracted fp32 ops);
grad_data within 1e-5 relative and bit-reproducible run to run."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import tf_ops as T

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))


def _dev():
    return torch.device("cuda:0")


def _run(data, wf, flags=0, go=None, need_data_grad=True):
    from dynamic_multiview_3d_b200 import functional as F
    d = torch.from_numpy(data).to(_dev()).requires_grad_(need_data_grad)
    w = torch.from_numpy(wf).to(_dev()).requires_grad_(True)
    out, idx, mask = F.resampler_debug(d.detach(), w.detach(), flags)
    res = {"out": out.cpu().numpy(), "idx": idx.cpu().numpy(), "mask": mask.cpu().numpy()}
    if go is not None:
        o2 = F._Resampler.apply(d, w, flags)
        assert torch.equal(o2.detach(), out)
        o2.backward(torch.from_numpy(go).to(_dev()))
        res["gw"] = w.grad.cpu().numpy()
        if need_data_grad:
            res["gd"] = d.grad.cpu().numpy()
    return res


def _check(data, warp, go, r, flags_add_grid_flow=None):
    fx, fy, cx, cy, mask = T.resampler_indices(data.shape, warp)
    assert np.array_equal(r["idx"], np.stack([fx, fy, cx, cy], -1)), "corner indices not bit-exact"
    assert np.array_equal(r["mask"], mask), "validity masks not bit-exact"
    ref = T.resampler(data, warp)
    assert np.array_equal(r["out"], ref), "max |diff| %g" % np.abs(r["out"] - ref).max()
    if go is not None:
        gd, gw = T.resampler_grad(data, warp, go)
        assert np.array_equal(r["gw"], gw), "grad_warp max |diff| %g" % np.abs(r["gw"] - gw).max()
        if "gd" in r:
            scale = max(np.abs(gd).max(), 1e-30)
            assert np.abs(r["gd"] - gd).max() <= 1e-5 * scale, np.abs(r["gd"] - gd).max() / scale


def test_golden_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "sampler_kat.npz"))
    r = _run(g["kat_img"], g["kat_pts"].reshape(1, -1, 2).copy())
    exp = g["kat_expect"]
    ok = ~np.isnan(exp)
    assert np.array_equal(r["out"][0, ok, 0], exp[ok])
    assert r["out"][0, -1, 0] == pytest.approx(1e-6, rel=0.05)
    for C in (1, 3, 4):
        data, warp, go = g["c%d_data" % C], g["c%d_warp" % C], g["c%d_go" % C]
        r = _run(data, warp, 0, go)
        assert np.array_equal(r["out"], g["c%d_out" % C])
        assert np.array_equal(r["idx"], g["c%d_idx" % C]) and np.array_equal(r["mask"], g["c%d_mask" % C])
        assert np.array_equal(r["gw"], g["c%d_gw" % C])
        assert np.allclose(r["gd"], g["c%d_gd" % C], rtol=1e-5, atol=1e-7)
    rq = _run(g["quirk_img"], np.zeros((1, 4, 4, 2), np.float32), flags=1)       # ADD_GRID, reference (Y,X) order
    assert np.array_equal(rq["out"], g["quirk_out"])
    assert np.array_equal(rq["out"][0, :, :, 0], g["quirk_img"][0, :, :, 0].T)
    rxy = _run(g["quirk_img"], np.zeros((1, 4, 4, 2), np.float32), flags=3)      # XY order: identity
    assert np.array_equal(rxy["out"], g["quirk_img"])


# W = 70: rows are not 16-byte multiples -> generic tile kernel; W = 72: the TMA kernel for C in {1,3,4}
@pytest.mark.parametrize("geom", [(45, 70, 37, 51), (45, 72, 37, 52)])
@pytest.mark.parametrize("C", [1, 2, 3, 4, 6])
@pytest.mark.parametrize("regime", ["jitter", "rotation", "absolute", "boundary"])
def test_parity_regimes(C, regime, geom):
    rng = np.random.default_rng((C * 7919 + len(regime) * 104729 + geom[1]) % 2**31)
    B, H, W = 3, geom[0], geom[1]                           # ragged: not multiples of the 32x32 tile
    Ho, Wo = (H, W) if regime != "absolute" else geom[2:]
    data = rng.random((B, H, W, C), dtype=np.float32)
    ii, jj = np.meshgrid(np.arange(Ho, dtype=np.float32), np.arange(Wo, dtype=np.float32), indexing="ij")
    if regime == "jitter":          # training regime at init: identity + U(-3,3), XY order so H != W works
        warp = np.stack([jj, ii], -1)[None] + rng.uniform(-3, 3, (B, Ho, Wo, 2)).astype(np.float32)
    elif regime == "rotation":      # test_resampler.py's 10 degree field
        c, s = np.cos(np.radians(10)), np.sin(np.radians(10))
        warp = np.broadcast_to(np.stack([jj * c + ii * s, -jj * s + ii * c], -1)[None], (B, Ho, Wo, 2))
    elif regime == "absolute":      # every predicate, worst-case locality (falls off the staged path)
        warp = rng.uniform(-2, [W + 1, H + 1], (B, Ho, Wo, 2))
    else:
        xs = np.array([-1, -1 + 1e-6, -0.5, 0, W - 1, W - 0.5, W - 1e-4, W, np.nan], np.float32)
        ys = np.array([-1, -1 + 1e-6, -0.5, 0, H - 1, H - 0.5, H - 1e-4, H, np.inf], np.float32)
        gx, gy = np.meshgrid(xs, ys)
        warp = np.zeros((B, Ho, Wo, 2), np.float32) + 5.3
        warp[:, :9, :9, 0] = gx
        warp[:, :9, :9, 1] = gy
    warp = np.ascontiguousarray(warp, np.float32)
    go = rng.standard_normal((B, Ho, Wo, C)).astype(np.float32)
    r = _run(data, warp, 0, go)
    _check(data, warp, go, r)


@pytest.mark.parametrize("order", ["ref_yx", "xy"])
def test_fused_grid_matches_warp_pts_layer(order):
    rng = np.random.default_rng(9)
    B, H, C = 2, 64, 3
    data = rng.random((B, H, H, C), dtype=np.float32)
    flow = rng.uniform(-3, 3, (B, H, H, 2)).astype(np.float32)
    go = rng.standard_normal((B, H, H, C)).astype(np.float32)
    warp = T.warp_pts_layer(flow) if order == "ref_yx" else (flow + T.coords(H, H, B)[..., ::-1]).astype(np.float32)
    r = _run(data, flow, 1 | (2 if order == "xy" else 0), go)
    _check(data, warp, go, r)
    # the package-level helpers agree with the kernel-fused grid
    from dynamic_multiview_3d_b200 import tf_utils as U
    f = torch.from_numpy(flow).cuda()
    if order == "ref_yx":
        assert np.array_equal(U.warp_pts_layer(f).cpu().numpy(), warp)
        out = U.resample_layer(torch.from_numpy(data).cuda(), U.warp_pts_layer(f))
        assert np.array_equal(out.cpu().numpy(), r["out"])


@pytest.mark.parametrize("C", [1, 3, 4])
@pytest.mark.parametrize("regime", ["jitter", "rotation", "far", "boundary"])
def test_reference_grid_on_nonsquare_geometry(C, regime):
    """The reference's (Y,X) grid fused into the kernel (flags = ADD_GRID) on a ragged, NON-square output over a
    non-square source (out[i,j] = bilinear(src, x = i + f0, y = j + f1)): forward and grad_warp bit-equal to the oracle."""
    rng = np.random.default_rng(C * 31 + len(regime))
    B, H, W, Ho, Wo = 3, 52, 76, 72, 44                     # W*C, Wo*C multiples of 4; 32x32 tiles ragged on both sides
    data = rng.random((B, H, W, C), dtype=np.float32)
    if regime == "jitter":
        flow = rng.uniform(-3, 3, (B, Ho, Wo, 2)).astype(np.float32)
    elif regime == "rotation":
        ii, jj = np.meshgrid(np.arange(Ho, dtype=np.float32), np.arange(Wo, dtype=np.float32), indexing="ij")
        c, s_ = np.cos(np.radians(10)), np.sin(np.radians(10))
        flow = np.broadcast_to(np.stack([ii * c + jj * s_ - ii, -ii * s_ + jj * c - jj], -1)[None], (B, Ho, Wo, 2)).astype(np.float32)
    elif regime == "far":                                    # tap boxes beyond the staged window: global gathers
        flow = rng.uniform(-40, 40, (B, Ho, Wo, 2)).astype(np.float32)
    else:
        flow = np.zeros((B, Ho, Wo, 2), np.float32)
        flow[:, :, :, 0] = rng.choice(np.array([-1.0, -1 + 1e-6, -0.5, 0.0, 0.5, 1e6, np.nan], np.float32), (B, Ho, Wo))
        flow[:, :, :, 1] = rng.choice(np.array([-1.0, -0.25, 0.0, 0.75, 3.0, -1e6], np.float32), (B, Ho, Wo))
    flow = np.ascontiguousarray(flow)
    go = rng.standard_normal((B, Ho, Wo, C)).astype(np.float32)
    r = _run(data, flow, 1, go, need_data_grad=False)
    ii, jj = np.meshgrid(np.arange(Ho, dtype=np.float32), np.arange(Wo, dtype=np.float32), indexing="ij")
    warp = (flow + np.stack([ii, jj], -1)[None]).astype(np.float32)              # tf_utils.py:44-52: channel 0 += row, 1 += column
    _check(data, warp, go, r)


def test_sample_list_layout_and_empty_validity():
    rng = np.random.default_rng(4)
    data = rng.random((2, 8, 8, 3), dtype=np.float32)
    warp = rng.uniform(-1, 8, (2, 3000, 2)).astype(np.float32)          # [B, N, 2] sample list
    go = rng.standard_normal((2, 3000, 3)).astype(np.float32)
    r = _run(data, warp, 0, go)
    assert r["out"].shape == (2, 3000, 3)
    _check(data, warp, go, r)
    far = np.full((1, 5, 7, 2), 1e6, np.float32)                         # nothing valid anywhere
    r2 = _run(data[:1], far, 0, np.ones((1, 5, 7, 3), np.float32))
    assert not r2["out"].any() and not r2["gw"].any() and not r2["gd"].any() and not r2["mask"].any()


def test_grad_data_is_bit_deterministic_and_contended():
    """Many outputs hitting the same source cells (minification): exercises the ranked
    same-cell serialisation; two runs must agree bit for bit."""
    rng = np.random.default_rng(5)
    B, H, C = 2, 96, 3
    data = rng.random((B, 16, 16, C), dtype=np.float32)
    ii, jj = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="ij")
    warp = np.ascontiguousarray(np.broadcast_to(np.stack([jj / 6.0, ii / 6.0], -1)[None], (B, H, H, 2)), np.float32)
    go = rng.standard_normal((B, H, H, C)).astype(np.float32)
    r1 = _run(data, warp, 0, go)
    r2 = _run(data, warp, 0, go)
    assert np.array_equal(r1["gd"], r2["gd"]) and np.array_equal(r1["gw"], r2["gw"])
    _check(data, warp, go, r1)


def test_full_size_properties():
    """BASELINE size (64 x 224^2 x 3): properties that need no oracle pass at this size --
    linearity in data, constant images stay constant inside, grad_data sums to the sum of
    in-range weights * grad_out (adjointness <grad_out, S(data)> == <grad_data, data>)."""
    from dynamic_multiview_3d_b200 import functional as F
    g = torch.Generator(device="cuda").manual_seed(0)
    B, H, C = 64, 224, 3
    data = torch.rand((B, H, H, C), device="cuda", generator=g)
    flow = (torch.rand((B, H, H, 2), device="cuda", generator=g) - 0.5) * 6
    out = F.flow_resampler(data, flow)
    out2 = F.flow_resampler(2.0 * data, flow)
    assert torch.equal(out2, 2.0 * out)                                   # exact: scaling by 2 commutes with rounding
    ones = F.flow_resampler(torch.ones_like(data), flow)
    inner = ones[:, 8:-8, 8:-8]
    assert float((inner - 1).abs().max()) < 1e-6
    d = data.clone().requires_grad_(True)
    f = flow.clone().requires_grad_(True)
    go = torch.randn((B, H, H, C), device="cuda", generator=g)
    o = F.flow_resampler(d, f)
    o.backward(go)
    lhs = float((go.double() * o.detach().double()).sum())
    rhs = float((d.grad.double() * data.double()).sum())
    assert abs(lhs - rhs) <= 1e-6 * abs(lhs) + 1e-3
    assert torch.isfinite(f.grad).all() and float(f.grad.abs().max()) > 0
    # a spot-check of 2 images against the oracle at full resolution
    sel = [0, 63]
    warp = T.warp_pts_layer(flow[sel].cpu().numpy())
    ref = T.resampler(data[sel].cpu().numpy(), warp)
    assert np.array_equal(out[sel].cpu().numpy(), ref)
    gd, gw = T.resampler_grad(data[sel].cpu().numpy(), warp, go[sel].cpu().numpy())
    assert np.array_equal(f.grad[sel].cpu().numpy(), gw)
    assert np.abs(d.grad[sel].cpu().numpy() - gd).max() <= 1e-5 * np.abs(gd).max()


def test_rectangle_fixture_on_gpu(golden_dir):
    """The reference's rectangle demo (test_resampler.py) through the CUDA path: same bytes as
    the oracle (committed SHA-256), rectangle rotated by 10 degrees."""
    import hashlib
    import make_golden as mg
    from dynamic_multiview_3d_b200 import functional as F
    g = np.load(os.path.join(golden_dir, "rectangle.npz"))
    H, W = int(g["shape"][0]), int(g["shape"][1])
    im = mg.rectangle_image(H, W, tuple(int(v) for v in g["bounds"]), g["fg"], g["bg"])
    warp = mg.rectangle_warp(H, W)
    out = F.resampler(torch.from_numpy(im[None].astype(np.float32)).cuda(), torch.from_numpy(warp).cuda())
    out8 = out.cpu().numpy().astype(np.uint8)[0]
    assert hashlib.sha256(out8.tobytes()).digest() == g["sha256"].tobytes()
    rect = np.all(out8 == g["fg"], axis=-1)
    assert abs(mg.principal_angle(rect) % 180.0 - 10.0) < 0.5


@pytest.mark.parametrize("C,mode,order", [(3, "l2", "ref_yx"), (3, "l1", "xy"), (4, "l2", "xy"), (1, "l2", "ref_yx")])
def test_fused_warp_loss_matches_the_three_separate_calls(C, mode, order):
    """dmv_sampler_loss_fused == sampler forward + fused loss + sampler backward: gen and d loss / d flow bit for bit
    (same arithmetic), the loss to summation-order accuracy; and all three against the NumPy oracle."""
    from dynamic_multiview_3d_b200 import functional as F
    rng = np.random.default_rng(C * 10 + len(mode))
    B, H = 3, 72                                             # 72 = ragged 32x32 tiling, rows 16-byte aligned for C = 3
    data = rng.random((B, H, H, C), dtype=np.float32)
    target = rng.random((B, H, H, C), dtype=np.float32)
    flow = rng.uniform(-3, 3, (B, H, H, 2)).astype(np.float32)
    flow[0, :4, :4] = 500.0                                  # invalid samples: gen = 0, no flow gradient
    wts = [1.0, 0.5, 2.0, 0.1][:C]
    inv = 1.0 / (B * H * H)
    d, t = torch.from_numpy(data).cuda(), torch.from_numpy(target).cuda()
    f1 = torch.from_numpy(flow).cuda().requires_grad_(True)
    gen1 = F.flow_resampler(d, f1, order)
    l1 = F.reconstruction_loss(gen1, t, mode, weights=wts, inv_count=inv)
    l1.backward()
    f2 = torch.from_numpy(flow).cuda().requires_grad_(True)
    assert F.warp_loss_supported(d, f2)
    l2, gen2 = F.flow_resample_loss(d, f2, t, mode, weights=wts, inv_count=inv, grid_order=order)
    l2.backward()
    assert torch.equal(gen1.detach(), gen2)
    assert torch.equal(f1.grad, f2.grad)
    assert float(l2) == pytest.approx(float(l1), rel=1e-6)
    grid = T.coords(H, H, B) if order == "ref_yx" else T.coords(H, H, B)[..., ::-1]
    ref = T.resampler(data, (flow + grid).astype(np.float32))
    assert np.array_equal(gen2.cpu().numpy(), ref)
    dd = ref.astype(np.float64) - target
    w = np.asarray(wts, np.float64)
    rl = ((dd * dd if mode == "l2" else np.abs(dd)) * w).sum() * inv
    assert float(l2) == pytest.approx(rl, rel=1e-5)
