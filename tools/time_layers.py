"""Times individual C-ABI layer calls at the BASELINE shapes (B=64, 224^2 graph): 20 back-to-back launches
between CUDA events, inputs > L2 only for the big layers (explanatory, not a bench number).

    python tools/time_layers.py [filter]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dynamic_multiview_3d_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
L = _lib.load()
B = int(os.environ.get("B", 64))
filt = sys.argv[1] if len(sys.argv) > 1 else ""
st = torch.cuda.current_stream().cuda_stream
bf = torch.bfloat16


def ws_for(n):
    return torch.empty(max(int(n), 256), dtype=torch.uint8, device=dev)


def timeit(name, fn, flop):
    if filt and filt not in name:
        return
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 20
    print("%-34s %8.1f us  %7.1f TFLOP/s" % (name, us, flop / us / 1e6), flush=True)


def conv(name, H, k, s, cin, cout, f32in=False):
    Ho = -(-H // s)
    x = torch.randn((B, H, H, cin), device=dev).to(torch.float32 if f32in else bf)
    xdt = 1 if f32in else 0
    w = (torch.randn((k, k, cin, cout), device=dev) * 0.05).to(bf)
    b = torch.zeros(cout, device=dev)
    y = torch.empty((B, Ho, Ho, cout), device=dev, dtype=bf)
    dy = torch.randn((B, Ho, Ho, cout), device=dev).to(bf)
    dx = torch.empty_like(x)
    dw = torch.empty((k, k, cin, cout), device=dev)
    ws = ws_for(max(L.dmv_conv_workspace_size(B, H, H, cin, cout, k, k, s), L.dmv_wgrad_workspace_size(B, H, H, cin, cout, k, k, s)))
    flop = 2.0 * B * Ho * Ho * k * k * cin * cout
    timeit(name + " fwd", lambda: _lib.call("dmv_conv2d_fwd", x.data_ptr(), xdt, w.data_ptr(), b.data_ptr(), y.data_ptr(), 0, B, H, H, cin, cout,
                                            k, k, s, 1, ws.data_ptr(), ws.numel(), 0, st), flop)
    if not f32in:
        timeit(name + " dgrad", lambda: _lib.call("dmv_conv2d_dgrad", dy.data_ptr(), w.data_ptr(), dx.data_ptr(), x.data_ptr(), 1, B, H, H, cin, cout, k, k, s,
                                                  ws.data_ptr(), ws.numel(), 0, st), flop)
    timeit(name + " wgrad", lambda: _lib.call("dmv_conv2d_wgrad", x.data_ptr(), xdt, dy.data_ptr(), dw.data_ptr(), None, B, H, H, cin, cout, k, k,
                                              s, ws.data_ptr(), ws.numel(), 0, st), flop)


def deconv(name, Ho, k, s, cin, cout, f32out=False):
    Hi = -(-Ho // s)
    x = torch.randn((B, Hi, Hi, cin), device=dev).to(bf)
    w = (torch.randn((k, k, cout, cin), device=dev) * 0.05).to(bf)
    y = torch.empty((B, Ho, Ho, cout), device=dev, dtype=torch.float32 if f32out else bf)
    dy = torch.randn((B, Ho, Ho, cout), device=dev).to(torch.float32 if f32out else bf)
    dx = torch.empty_like(x)
    dw = torch.empty((k, k, cout, cin), device=dev)
    ws = ws_for(max(L.dmv_conv_workspace_size(B, Ho, Ho, cout, cin, k, k, s), L.dmv_wgrad_workspace_size(B, Ho, Ho, cout, cin, k, k, s)))
    flop = 2.0 * B * Hi * Hi * k * k * cin * cout
    ydt = 1 if f32out else 0
    timeit(name + " fwd", lambda: _lib.call("dmv_deconv2d_fwd", x.data_ptr(), w.data_ptr(), y.data_ptr(), ydt, B, Ho, Ho, cin, cout, k, k, s, 0,
                                            ws.data_ptr(), ws.numel(), 0, st), flop)
    timeit(name + " dgrad", lambda: _lib.call("dmv_deconv2d_dgrad", dy.data_ptr(), ydt, w.data_ptr(), dx.data_ptr(), x.data_ptr(), 1, B, Ho, Ho, cin, cout, k, k, s,
                                              ws.data_ptr(), ws.numel(), 0, st), flop)
    timeit(name + " wgrad", lambda: _lib.call("dmv_deconv2d_wgrad", x.data_ptr(), dy.data_ptr(), ydt, dw.data_ptr(), B, Ho, Ho, cin, cout, k, k, s,
                                              ws.data_ptr(), ws.numel(), 0, st), flop)


def linear(name, K, N):
    M = B
    x = torch.randn((M, K), device=dev).to(bf)
    w = (torch.randn((K, N), device=dev) * 0.02).to(bf)
    b = torch.zeros(N, device=dev)
    y = torch.empty((M, N), device=dev, dtype=bf)
    dy = torch.randn((M, N), device=dev).to(bf)
    dx = torch.empty_like(x)
    dw = torch.empty((K, N), device=dev)
    ws = ws_for(max(L.dmv_conv_workspace_size(M, 1, 1, K, N, 1, 1, 1), L.dmv_wgrad_workspace_size(M, 1, 1, K, N, 1, 1, 1)))
    flop = 2.0 * M * K * N
    nb = K * N
    timeit(name + " fwd", lambda: _lib.call("dmv_linear_fwd", x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, K, N, 1, ws.data_ptr(),
                                            ws.numel(), 0, st), flop)
    timeit(name + " dgrad", lambda: _lib.call("dmv_linear_dgrad", dy.data_ptr(), w.data_ptr(), dx.data_ptr(), x.data_ptr(), 1, M, K, N, ws.data_ptr(), ws.numel(),
                                              0, st), flop)
    timeit(name + " wgrad", lambda: _lib.call("dmv_linear_wgrad", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), None, M, K, N, ws.data_ptr(),
                                              ws.numel(), 0, st), flop)
    if not filt or filt in name:
        print("    (%s: weight stream %.1f MB bf16 -> %.1f us at 6.5 TB/s; dW %.1f MB fp32 -> %.1f us)" % (
            name, nb * 2 / 1e6, nb * 2 / 6.5e6, nb * 4 / 1e6, nb * 4 / 6.5e6))


linear("fc1 lin 12544>4096", 12544, 4096)
linear("a3 lin 4160>4096", 4160, 4096)
linear("a4 lin 4096>4096", 4096, 4096)
linear("a5 lin 4096>12544", 4096, 12544)
conv("e0 c5s2 224 3>32", 224, 5, 2, 3, 32, True)
conv("e0_0 c5s1 112 32>32", 112, 5, 1, 32, 32)
conv("e1 c5s2 112 32>32", 112, 5, 2, 32, 32)
conv("e1_0 c5s1 56 32>32", 56, 5, 1, 32, 32)
conv("e2 c5s2 56 32>64", 56, 5, 2, 32, 64)
conv("e2_0 c5s1 28 64>64", 28, 5, 1, 64, 64)
conv("e3 c3s2 28 64>128", 28, 3, 2, 64, 128)
conv("e3_0 c3s1 14 128>128", 14, 3, 1, 128, 128)
conv("e4 c3s2 14 128>256", 14, 3, 2, 128, 256)
conv("e4_0 c3s1 7 256>256", 7, 3, 1, 256, 256)
conv("d2_0 c5s1 56 32>64", 56, 5, 1, 32, 64)
deconv("d4 d3s2 14 256>128", 14, 3, 2, 256, 128)
deconv("d3 d3s2 28 128>64", 28, 3, 2, 128, 64)
deconv("d2 d5s2 56 64>32", 56, 5, 2, 64, 32)
deconv("d1 d5s2 112 64>32", 112, 5, 2, 64, 32)
deconv("flow d5s2 224 32>2", 224, 5, 2, 32, 2, True)
