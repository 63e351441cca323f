"""Standalone timing of dmv_linear_wgrad_adam (csrc/fc_adam.cu) on the four FC shapes of the 224^2 graph at batch 64,
next to the two calls it replaces (dmv_linear_wgrad + dmv_adam_multi).  One kernel variant per process
(DMV_FC_ADAM_VARIANT); CUDA events around 10 back-to-back launches; each layer's parameter state is 0.2 - 0.7 GB >> L2."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamic_multiview_3d_b200 import _lib  # noqa: E402

SHAPES = [("fc1", 64, 12544, 4096), ("a3", 64, 4160, 4096), ("a4", 64, 4096, 4096), ("a5", 64, 4096, 12544)]


def main():
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    state = torch.tensor([0.9, 0.999, 1e-4, 1.0], device=dev)
    variant = os.environ.get("DMV_FC_ADAM_VARIANT", "default")
    tot_f = tot_s = 0.0
    for name, M, K, N in SHAPES:
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        dy = (torch.randn(M, N, device=dev) * 1e-3).to(torch.bfloat16)
        th = torch.randn(K, N, device=dev) * 0.02
        m, v = torch.zeros_like(th), torch.zeros_like(th)
        hf = th.to(torch.bfloat16)
        dw = torch.empty_like(th)
        nws = _lib.load().dmv_wgrad_workspace_size(M, 1, 1, K, N, 1, 1, 1)
        ws = torch.empty(int(nws), dtype=torch.uint8, device=dev)
        vp = C.c_void_p * 1

        def fused():
            _lib.call("dmv_linear_wgrad_adam", x.data_ptr(), dy.data_ptr(), th.data_ptr(), m.data_ptr(), v.data_ptr(), hf.data_ptr(), None,
                      M, K, N, state.data_ptr(), 0.9, 0.999, 1e-8, 1.0, st)

        def separate():
            _lib.call("dmv_linear_wgrad", x.data_ptr(), dy.data_ptr(), dw.data_ptr(), None, M, K, N, ws.data_ptr(), ws.numel(), 0, st)
            _lib.call("dmv_adam_multi", vp(th.data_ptr()), vp(dw.data_ptr()), vp(m.data_ptr()), vp(v.data_ptr()), vp(hf.data_ptr()),
                      (C.c_longlong * 1)(K * N), 1, state.data_ptr(), 0.9, 0.999, 1e-8, 1.0, st)

        res = {}
        for tag, fn in (("fused", fused), ("separate", separate)):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[tag] = e0.elapsed_time(e1) / 10
        tot_f += res["fused"]
        tot_s += res["separate"]
        print("variant %s  %-4s %5dx%-5d  fused %7.1f us = %6.0f GB/s (26 B/param)   wgrad + adam %7.1f us = %6.0f GB/s (34 B/param)"
              % (variant, name, K, N, 1e3 * res["fused"], 26.0 * K * N / res["fused"] / 1e6, 1e3 * res["separate"], 34.0 * K * N / res["separate"] / 1e6))
        del x, dy, th, m, v, hf, dw, ws
    print("variant %s  all four: fused %.1f us, separate %.1f us" % (variant, 1e3 * tot_f, 1e3 * tot_s))


if __name__ == "__main__":
    main()
