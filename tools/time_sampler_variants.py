"""Times sampler fwd / grad_flow of every tools/variants/*.so at 64x224^2x3 and checks they agree bit for bit."""
import ctypes as C
import glob
import os
import sys

import torch

B, H, Cc = 64, 224, 3
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
import math
import sys
REGIME = sys.argv[1] if len(sys.argv) > 1 else "jitter"
ii, jj = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32), torch.arange(H, device=dev, dtype=torch.float32), indexing="ij")
sets = []
for k in range(4):
    if REGIME == "jitter":      # U(-3,3) around identity
        flow = (torch.rand((B, H, H, 2), device=dev, generator=g) - 0.5) * 6
    else:                       # smooth: rotation by (8 + k) degrees about the centre, as flow relative to the (Y,X) grid
        a = math.radians(8.0 + k)
        c0 = (H - 1) / 2
        u = math.cos(a) * (ii - c0) - math.sin(a) * (jj - c0) + c0
        v = math.sin(a) * (ii - c0) + math.cos(a) * (jj - c0) + c0
        flow = torch.stack([u - ii, v - jj], -1).unsqueeze(0).repeat(B, 1, 1, 1).contiguous()
    sets.append((torch.rand((B, H, H, Cc), device=dev, generator=g), flow, torch.randn((B, H, H, Cc), device=dev, generator=g)))
print("regime", REGIME)
out = torch.empty((B, H, H, Cc), device=dev)
gw = torch.empty((B, H, H, 2), device=dev)
ref = None
vp, i = C.c_void_p, C.c_int
for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "variants", "*.so"))):
    lib = C.CDLL(path)
    lib.dmv_sampler_fwd.argtypes = [vp] * 5 + [i] * 6 + [C.c_uint, vp]
    lib.dmv_sampler_bwd.argtypes = [vp] * 5 + [i] * 6 + [C.c_uint, vp, C.c_size_t, vp]
    st = torch.cuda.current_stream().cuda_stream

    def fwd(d, f, go):
        assert lib.dmv_sampler_fwd(d.data_ptr(), f.data_ptr(), out.data_ptr(), None, None, B, H, H, Cc, H, H, 1, st) == 0

    def bwd(d, f, go):
        assert lib.dmv_sampler_bwd(d.data_ptr(), f.data_ptr(), go.data_ptr(), None, gw.data_ptr(), B, H, H, Cc, H, H, 1, None, 0, st) == 0

    res = []
    for fn in (fwd, bwd):
        for k in range(3):
            fn(*sets[k % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(40):
            fn(*sets[k % 4])
        e1.record()
        torch.cuda.synchronize()
        res.append(1e3 * e0.elapsed_time(e1) / 40)
    fwd(*sets[0]); bwd(*sets[0])
    torch.cuda.synchronize()
    sig = (out.clone(), gw.clone())
    if ref is None:
        ref = sig
    same = torch.equal(sig[0], ref[0]) and torch.equal(sig[1], ref[1])
    print("%-16s fwd %6.1f us (%.0f GB/s)   grad_flow %6.1f us (%.0f GB/s)   bit-equal %s" % (
        os.path.basename(path)[:-3], res[0], B * H * H * 32 / res[0] / 1e3, res[1], B * H * H * 40 / res[1] / 1e3, same), flush=True)
