"""Runs the halo-descriptor probe on the GPU box and prints which (row offset, base_offset) settings give
D[m] == x[bh + r, bw + s] for m = bh*8 + bw."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(HERE, "libprobe.so")
if not os.path.exists(lib_path):
    csrc = os.path.join(HERE, "..", "..", "dynamic_multiview_3d_b200", "csrc")
    subprocess.run(["nvcc", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
                    "-shared", "-o", lib_path, os.path.join(HERE, "probe.cu"), os.path.join(csrc, "api.cu")], check=True)
lib = ctypes.CDLL(lib_path)
lib.probe_halo.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 4 + [ctypes.c_void_p]

for C in (64, 32):
    rb = C * 2
    x = torch.arange(24 * 16, device="cuda", dtype=torch.float32).reshape(24, 16, 1).repeat(1, 1, C)
    x = (x + torch.arange(C, device="cuda").reshape(1, 1, C) * 0.001 * 0).to(torch.bfloat16)      # value = pixel index (exact in bf16 < 256?)
    x = (torch.arange(24 * 16, device="cuda").reshape(24, 16, 1) % 251).float().repeat(1, 1, C)
    x = x + (torch.arange(C, device="cuda").reshape(1, 1, C) % 4).float() * 0.25                    # distinguishes channels mod 4
    xb = x.to(torch.bfloat16).contiguous()
    ident = torch.eye(C, device="cuda").to(torch.bfloat16).contiguous()
    out = torch.zeros((128, C), device="cuda", dtype=torch.float32)
    print("==== C=%d (row bytes %d)" % (C, rb))
    for r in (0, 1, 2, 4):
        for s in (0, 1, 2, 3, 4):
            row_off = r * 16 + s
            exp = xb[r:r + 16, s:s + 8].reshape(128, C).float()
            res = []
            start_phase = ((row_off * rb) >> 7) & 7
            for bo in sorted({0, start_phase, s & 7, (s >> 1) & 3}):
                out.zero_()
                rc = lib.probe_halo(xb.data_ptr(), ident.data_ptr(), out.data_ptr(), C, row_off, bo, 16 * rb, None)
                torch.cuda.synchronize()
                ok = bool(torch.equal(out, exp))
                nbad = int((out != exp).any(dim=1).sum())
                res.append("bo=%d:%s(%d bad rows)" % (bo, "OK" if ok else "X", nbad))
            print("r=%d s=%d row_off=%3d start_phase=%d  %s" % (r, s, row_off, start_phase, "  ".join(res)))
