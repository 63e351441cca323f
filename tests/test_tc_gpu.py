"""GPU tests of the tcgen05/TMEM/TMA implicit-GEMM path: it must actually run (tensor-core
launch counter), agree with the oracle on small cases and with the SIMT kernels at the
BASELINE layer sizes (224^2 graph, batch 8)."""
import numpy as np
import pytest
import torch

from oracle import tf_ops as T

pytestmark = pytest.mark.gpu


def _store():
    from dynamic_multiview_3d_b200.variables import VariableStore
    return VariableStore(torch.device("cuda:0"))


def _var(store, name, arr):
    v = store.get(name, arr.shape, "zeros")
    v.master.copy_(torch.from_numpy(np.ascontiguousarray(arr)).cuda())
    store._cast(v.master, v.half, v.numel)
    return v


def bf16_round(x):
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


SMALL = [  # kind, k, s, H(big side), cin, cout, B
    ("conv", 5, 1, 16, 32, 32, 3), ("conv", 5, 2, 16, 32, 64, 3), ("conv", 3, 1, 7, 256, 256, 5), ("conv", 3, 2, 14, 128, 256, 3),
    ("conv", 3, 1, 14, 128, 128, 2), ("conv", 5, 1, 28, 64, 64, 2), ("conv", 3, 2, 28, 64, 128, 2), ("conv", 5, 1, 24, 32, 64, 1),
    ("deconv", 3, 2, 14, 256, 128, 3), ("deconv", 3, 2, 28, 128, 64, 2), ("deconv", 5, 2, 16, 64, 32, 3), ("deconv", 5, 2, 32, 32, 2, 2),
    ("deconv", 3, 1, 16, 32, 2, 2),
    # halo weight-gradient kernel: ragged M tiles (3 taps per row, 128/CB = 4 or 2 per MMA), ragged image tiling
    ("conv", 3, 1, 20, 32, 32, 2), ("conv", 3, 1, 18, 64, 64, 2), ("conv", 5, 1, 22, 64, 32, 2), ("conv", 4, 1, 16, 32, 32, 1),
    # sub-pixel form of the stride-2 G problems (deconv fwd / conv dgrad as one stride-1 halo launch + depth-to-space)
    ("deconv", 3, 2, 32, 64, 32, 2), ("conv", 3, 2, 36, 32, 32, 2), ("deconv", 5, 2, 34, 32, 16, 1), ("conv", 5, 2, 36, 32, 64, 1),
    ("deconv", 5, 2, 36, 64, 64, 1),
    # stride-2 F problems with 32 source channels on the halo kernel (two row-parity planes, column parity folded into K)
    ("conv", 5, 2, 36, 32, 32, 2), ("conv", 3, 2, 34, 32, 64, 1), ("deconv", 5, 2, 36, 32, 32, 2), ("deconv", 3, 2, 40, 32, 32, 1),
]


@pytest.mark.parametrize("kind,k,s,H,cin,cout,B", SMALL)
def test_tc_small_vs_oracle(kind, k, s, H, cin, cout, B):
    from dynamic_multiview_3d_b200 import _lib, functional as F
    rng = np.random.default_rng(k * 1000 + s * 100 + cin + cout + H)
    st = _store()
    n0 = _lib.tc_launch_count()
    if kind == "conv":
        x = bf16_round(rng.standard_normal((B, H, H, cin)))
        w = bf16_round(rng.standard_normal((k, k, cin, cout)) * T.conv_stddev(k, k, cin))
        b = rng.standard_normal(cout).astype(np.float32)
        y = T.conv2d_same(x, w, b, s, s)
        gy = bf16_round(rng.standard_normal(y.shape))
        gx, gw, gb = T.conv2d_same_grads(x, w, gy, s, s)
        wv, bv = _var(st, "w", w), _var(st, "b", b)
        xt = torch.from_numpy(x).cuda().to(torch.bfloat16).requires_grad_(True)
        yo = F.conv2d(xt, wv, bv, s, None, "auto", torch.float32)
        assert _lib.tc_launch_count() == n0 + 1, "forward did not take the tensor-core path"
        assert _rel(yo.detach().cpu().numpy(), y) < 1e-4
        yb = F.conv2d(xt, wv, bv, s, "lrelu", "auto")
        assert _rel(yb.detach().float().cpu().numpy(), T.lrelu(y)) < 1e-2
        yo2 = F.conv2d(xt, wv, bv, s, None, "auto")
        n1 = _lib.tc_launch_count()
        yo2.backward(torch.from_numpy(gy).cuda().to(torch.bfloat16))
        assert _lib.tc_launch_count() == n1 + 2, "dgrad / wgrad did not take the tensor-core path"
    else:
        h = -(-H // s)
        x = bf16_round(rng.standard_normal((B, h, h, cin)))
        w = bf16_round(rng.standard_normal((k, k, cout, cin)) * T.deconv_stddev(k, k, cin, s, s))
        y = T.conv2d_transpose_same(x, w, (B, H, H, cout), s, s)
        gy = bf16_round(rng.standard_normal(y.shape))
        gx, gw = T.conv2d_transpose_same_grads(x, w, gy, s, s)
        wv = _var(st, "w", w)
        xt = torch.from_numpy(x).cuda().to(torch.bfloat16).requires_grad_(True)
        yo = F.deconv2d(xt, wv, (H, H), s, None, "auto", torch.float32)
        assert _lib.tc_launch_count() == n0 + 1, "forward did not take the tensor-core path"
        assert _rel(yo.detach().cpu().numpy(), y) < 1e-4
        yo2 = F.deconv2d(xt, wv, (H, H), s, None, "auto")
        n1 = _lib.tc_launch_count()
        yo2.backward(torch.from_numpy(gy).cuda().to(torch.bfloat16))
        if cout % 32 == 0:
            assert _lib.tc_launch_count() == n1 + 2, "dgrad / wgrad did not take the tensor-core path"
    assert _rel(xt.grad.float().cpu().numpy(), gx) < 1e-2
    assert _rel(wv.grad.cpu().numpy(), gw) < 2e-4


FULL = [  # the 224^2 graph's tensor-core layers (appearance_flow_model.py:89-125), batch 8
    ("conv", 5, 1, 112, 32, 32), ("conv", 5, 2, 112, 32, 32), ("conv", 5, 1, 56, 32, 32), ("conv", 5, 2, 56, 32, 64),
    ("conv", 5, 1, 28, 64, 64), ("conv", 3, 2, 28, 64, 128), ("conv", 3, 1, 14, 128, 128), ("conv", 3, 2, 14, 128, 256),
    ("conv", 3, 1, 7, 256, 256), ("conv", 5, 1, 56, 32, 64),
    ("deconv", 3, 2, 14, 256, 128), ("deconv", 3, 2, 28, 128, 64), ("deconv", 5, 2, 56, 64, 32), ("deconv", 5, 2, 112, 64, 32),
    ("deconv", 5, 2, 224, 32, 2),
]


@pytest.mark.parametrize("kind,k,s,H,cin,cout", FULL)
def test_tc_full_size_vs_simt(kind, k, s, H, cin, cout):
    """BASELINE layer sizes: the oracle is too slow here, so the tensor-core result is compared with the
    (oracle-verified) SIMT kernels on the same bf16 inputs -- both accumulate in fp32."""
    from dynamic_multiview_3d_b200 import _lib, functional as F
    B = 8
    g = torch.Generator(device="cuda").manual_seed(k + s + H + cin)
    st = _store()
    wgs = {}
    if kind == "conv":
        x = torch.randn((B, H, H, cin), device="cuda", generator=g).to(torch.bfloat16)
        wv = _var(st, "w", (np.random.default_rng(1).standard_normal((k, k, cin, cout)) * T.conv_stddev(k, k, cin)).astype(np.float32))
        bv = _var(st, "b", np.random.default_rng(2).standard_normal(cout).astype(np.float32))
        outs, grads = {}, {}
        for algo in ("simt", "auto"):
            xt = x.clone().requires_grad_(True)
            n0 = _lib.tc_launch_count()
            y = F.conv2d(xt, wv, bv, s, "lrelu", algo, torch.float32)
            if algo == "auto":
                assert _lib.tc_launch_count() == n0 + 1
            gy = torch.randn(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
            yb = F.conv2d(xt, wv, bv, s, None, algo)
            yb.backward(gy.to(torch.bfloat16))
            outs[algo], grads[algo] = y.detach(), xt.grad.float()
            wgs[algo] = wv.grad.clone()
    else:
        h = -(-H // s)
        x = torch.randn((B, h, h, cin), device="cuda", generator=g).to(torch.bfloat16)
        wv = _var(st, "w", (np.random.default_rng(1).standard_normal((k, k, cout, cin)) * T.deconv_stddev(k, k, cin, s, s)).astype(np.float32))
        outs, grads = {}, {}
        for algo in ("simt", "auto"):
            xt = x.clone().requires_grad_(True)
            n0 = _lib.tc_launch_count()
            y = F.deconv2d(xt, wv, (H, H), s, "lrelu" if cout > 2 else None, algo, torch.float32)
            if algo == "auto":
                assert _lib.tc_launch_count() == n0 + 1
            gy = torch.randn(y.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
            yb = F.deconv2d(xt, wv, (H, H), s, None, algo)
            yb.backward(gy.to(torch.bfloat16))
            outs[algo], grads[algo] = y.detach(), xt.grad.float()
            wgs[algo] = wv.grad.clone()
    dw = (wgs["auto"] - wgs["simt"]).abs().max() / wgs["simt"].abs().max()
    assert float(dw) < 1e-3, float(dw)           # fp32 sums over up to 100k pixels in different orders
    d = (outs["auto"] - outs["simt"]).abs().max() / outs["simt"].abs().max()
    assert float(d) < 1e-4, float(d)
    dg = (grads["auto"] - grads["simt"]).abs().max() / grads["simt"].abs().max()
    assert float(dg) < 1e-2, float(dg)           # bf16 outputs: one rounding apart at most


@pytest.mark.parametrize("M,K,N", [(64, 12544, 4096), (64, 4160, 4096), (64, 4096, 12544), (64, 64, 64), (8, 4096, 256), (3, 64, 64)])
def test_tc_linear_wgrad(M, K, N):
    from dynamic_multiview_3d_b200 import _lib, functional as F
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    x = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
    gy = torch.randn((M, N), device="cuda", generator=g).to(torch.bfloat16)
    st = _store()
    mv = _var(st, "Matrix", np.zeros((K, N), np.float32))
    bv = _var(st, "b", np.zeros((N,), np.float32))
    n0 = _lib.tc_launch_count()
    xt = x.clone().requires_grad_(True)
    F.linear(xt, mv, bv, None, "auto").backward(gy)
    assert _lib.tc_launch_count() > n0, "linear wgrad did not take the tensor-core path"
    ref = x.float().t() @ gy.float()
    d = (mv.grad - ref).abs().max() / ref.abs().max()
    assert float(d) < 1e-5, float(d)
    assert float((bv.grad - gy.float().sum(0)).abs().max()) < 1e-3


@pytest.mark.parametrize("M,K,N", [(64, 12544, 4096), (64, 4160, 4096), (64, 4096, 12544), (64, 64, 64), (8, 4096, 256), (3, 64, 64),
                                   (130, 96, 72)])
def test_tc_linear_fwd_dgrad(M, K, N):
    from dynamic_multiview_3d_b200 import _lib, functional as F
    rng = np.random.default_rng(M + K + N)
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    x = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
    gy = torch.randn((M, N), device="cuda", generator=g).to(torch.bfloat16)
    w = (rng.standard_normal((K, N)) * T.linear_stddev(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    st = _store()
    mv, bv = _var(st, "Matrix", w), _var(st, "b", b)
    wh = mv.half.float()
    n0 = _lib.tc_launch_count()
    xt = x.clone().requires_grad_(True)
    y = F.linear(xt, mv, bv, "lrelu", "auto")
    assert _lib.tc_launch_count() == n0 + 1, "linear forward did not take the tensor-core path"
    pre = x.float() @ wh + torch.from_numpy(b).cuda()
    ref = 0.6 * pre + 0.4 * pre.abs()
    assert float((y.float() - ref).abs().max() / ref.abs().max()) < 1e-2
    y2 = F.linear(xt, mv, bv, None, "auto")
    n1 = _lib.tc_launch_count()
    y2.backward(gy)
    if N % 32 == 0 and K % 32 == 0:
        assert _lib.tc_launch_count() >= n1 + 2, "linear dgrad / wgrad did not take the tensor-core path"
    gref = gy.float() @ wh.t()
    assert float((xt.grad.float() - gref).abs().max() / gref.abs().max()) < 1e-2


DACT = [  # kind, k, s, H (big side), cin, cout : every input-gradient kernel form of the graph (halo, sub-pixel C=32 and generic,
          # parity planes, per-tap igemm, thin space-to-depth head) at batch 8
    ("conv", 5, 1, 112, 32, 32), ("conv", 5, 2, 112, 32, 32), ("conv", 5, 2, 56, 32, 64), ("conv", 3, 2, 28, 64, 128),
    ("conv", 3, 1, 14, 128, 128), ("conv", 3, 1, 7, 256, 256), ("conv", 5, 1, 56, 32, 64),
    ("deconv", 3, 2, 14, 256, 128), ("deconv", 5, 2, 56, 64, 32), ("deconv", 5, 2, 112, 64, 32), ("deconv", 5, 2, 224, 32, 2),
]


@pytest.mark.parametrize("kind,k,s,H,cin,cout", DACT)
def test_dgrad_with_fused_activation_derivative(kind, k, s, H, cin, cout):
    """include/dmv3d.h y_in / act_in: dgrad(..., y_in, lrelu) == act_bwd(dgrad(...), y_in) up to the rounding it saves
    (the factor is applied to the fp32 accumulator): bit-identical where the slope is 1, within one bf16 ulp elsewhere."""
    from dynamic_multiview_3d_b200 import _lib
    L = _lib.load()
    B = 8
    g = torch.Generator(device="cuda").manual_seed(k + s + H + cin + cout)
    st = torch.cuda.current_stream().cuda_stream
    bf = torch.bfloat16
    if kind == "conv":
        xs, ys = (B, H, H, cin), (B, -(-H // s), -(-H // s), cout)
        w = (torch.randn((k, k, cin, cout), device="cuda", generator=g) * T.conv_stddev(k, k, cin)).to(bf)
    else:
        xs, ys = (B, -(-H // s), -(-H // s), cin), (B, H, H, cout)
        w = (torch.randn((k, k, cout, cin), device="cuda", generator=g) * T.deconv_stddev(k, k, cin, s, s)).to(bf)
    y_in = torch.randn(xs, device="cuda", generator=g).to(bf)                       # the layer's input = the producer's lrelu output
    dy = torch.randn(ys, device="cuda", generator=g).to(bf if cout >= 8 else torch.float32)
    ws = torch.empty(max(L.dmv_conv_workspace_size(B, H, H, cin if kind == "conv" else cout, cout if kind == "conv" else cin, k, k, s), 256) + (64 << 20),
                     dtype=torch.uint8, device="cuda")
    outs = []
    for fused in (False, True):
        dx = torch.empty(xs, dtype=bf, device="cuda")
        yp, act = (y_in.data_ptr(), 1) if fused else (None, 0)
        if kind == "conv":
            _lib.call("dmv_conv2d_dgrad", dy.data_ptr(), w.data_ptr(), dx.data_ptr(), yp, act, B, H, H, cin, cout, k, k, s, ws.data_ptr(), ws.numel(), 0, st)
        else:
            _lib.call("dmv_deconv2d_dgrad", dy.data_ptr(), 0 if dy.dtype == bf else 1, w.data_ptr(), dx.data_ptr(), yp, act, B, H, H, cin, cout, k, k, s,
                      ws.data_ptr(), ws.numel(), 0, st)
        if not fused:
            _lib.call("dmv_act_bwd", dx.data_ptr(), y_in.data_ptr(), dx.data_ptr(), 0, dx.numel(), 1, st)
        outs.append(dx.float())
    two_pass, fused = outs
    pos = y_in.float() > 0
    assert torch.equal(two_pass[pos], fused[pos])
    assert float(((two_pass - fused).abs() - 2.0 ** -7 * fused.abs()).max()) <= 0.0
    assert float(fused.abs().max()) > 0


@pytest.mark.parametrize("M,K,N", [(64, 12544, 4096), (64, 4096, 12544), (64, 4160, 4096), (64, 64, 64)])
def test_linear_dgrad_with_fused_activation_derivative(M, K, N):
    """The split-K linear input gradient applies the factor in its finish pass."""
    from dynamic_multiview_3d_b200 import _lib
    L = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    st = torch.cuda.current_stream().cuda_stream
    bf = torch.bfloat16
    w = (torch.randn((K, N), device="cuda", generator=g) * T.linear_stddev(K)).to(bf)
    y_in = torch.randn((M, K), device="cuda", generator=g).to(bf)
    dy = torch.randn((M, N), device="cuda", generator=g).to(bf)
    ws = torch.empty(max(L.dmv_conv_workspace_size(M, 1, 1, K, N, 1, 1, 1), 256), dtype=torch.uint8, device="cuda")
    outs = []
    for fused in (False, True):
        dx = torch.empty((M, K), dtype=bf, device="cuda")
        _lib.call("dmv_linear_dgrad", dy.data_ptr(), w.data_ptr(), dx.data_ptr(), y_in.data_ptr() if fused else None, 1 if fused else 0, M, K, N,
                  ws.data_ptr(), ws.numel(), 0, st)
        if not fused:
            _lib.call("dmv_act_bwd", dx.data_ptr(), y_in.data_ptr(), dx.data_ptr(), 0, dx.numel(), 1, st)
        outs.append(dx.float())
    two_pass, fused = outs
    pos = y_in.float() > 0
    assert torch.equal(two_pass[pos], fused[pos])
    assert float(((two_pass - fused).abs() - 2.0 ** -7 * fused.abs()).max()) <= 0.0
    ref = (dy.float() @ w.float().t()) * torch.where(pos, 1.0, 0.2)
    assert float((fused - ref).abs().max() / ref.abs().max()) < 1e-2


@pytest.mark.parametrize("H,B,ct", [(32, 2, 3), (32, 3, 2), (64, 2, 4), (224, 2, 3), (224, 2, 2), (96, 1, 1)])
def test_thin_layers_single_kernel_form(H, B, ct):
    """csrc/thin_mma.cu: the thin stride-2 5x5 layers (e0: ct -> 32 conv on fp32 pixels; flow head: 32 -> ct deconv with fp32
    output and fp32 output gradient) as one mma.sync kernel each, against the oracle: forward of both, the flow head's input
    gradient (the same kernel as e0's forward) and both weight gradients (window staged once, contraction over pixels)."""
    from dynamic_multiview_3d_b200 import _lib, functional as F
    rng = np.random.default_rng(H * 10 + B + ct)
    st = _store()
    # --- conv side (e0): fp32 input in [0, 1] (rounded to bf16 inside the kernel: compare against the oracle on rounded pixels)
    x = rng.random((B, H, H, ct), dtype=np.float32)
    w = bf16_round(rng.standard_normal((5, 5, ct, 32)) * T.conv_stddev(5, 5, ct))
    b = rng.standard_normal(32).astype(np.float32) * 0.1
    y = T.conv2d_same(bf16_round(x), w, b, 2, 2)
    gy = bf16_round(rng.standard_normal(y.shape))
    _, gw, gb = T.conv2d_same_grads(bf16_round(x), w, gy, 2, 2)
    wv, bv = _var(st, "e0/w", w), _var(st, "e0/b", b)
    assert _lib.load().dmv_thin_s2d_size(B, H, H, ct, 32, 5, 5, 2) == 0           # no prep tensor on this path
    xt = torch.from_numpy(x).cuda()
    n0 = _lib.launch_count()
    yo = F.conv2d(xt, wv, bv, 2, "lrelu", "auto")
    assert _lib.launch_count() == n0 + 1, "e0 forward is one kernel"
    assert _rel(yo.detach().float().cpu().numpy(), T.lrelu(y)) < 1e-2
    yl = F.conv2d(xt, wv, bv, 2, None, "auto")
    yl.backward(torch.from_numpy(gy).cuda().to(torch.bfloat16))
    assert _rel(wv.grad.cpu().numpy(), gw) < 2e-4
    assert _rel(bv.grad.cpu().numpy(), gb) < 1e-4
    # --- deconv side (flow head): bf16 input, fp32 output, fp32 output gradient
    h = H // 2
    xd = bf16_round(rng.standard_normal((B, h, h, 32)))
    wd = bf16_round(rng.standard_normal((5, 5, ct, 32)) * T.deconv_stddev(5, 5, 32, 2, 2))
    yd = T.conv2d_transpose_same(xd, wd, (B, H, H, ct), 2, 2)
    gyd = bf16_round(rng.standard_normal(yd.shape))
    gxd, gwd = T.conv2d_transpose_same_grads(xd, wd, gyd, 2, 2)
    wdv = _var(st, "flow/w", wd)
    xdt = torch.from_numpy(xd).cuda().to(torch.bfloat16).requires_grad_(True)
    n0 = _lib.launch_count()
    ydo = F.deconv2d(xdt, wdv, (H, H), 2, None, "auto", torch.float32)
    assert _lib.launch_count() == n0 + 1, "flow head forward is one kernel"
    assert _rel(ydo.detach().cpu().numpy(), yd) < 1e-4
    ydo.backward(torch.from_numpy(gyd).cuda())
    assert _rel(xdt.grad.float().cpu().numpy(), gxd) < 1e-2
    assert _rel(wdv.grad.cpu().numpy(), gwd) < 2e-4
