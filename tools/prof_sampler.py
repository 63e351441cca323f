"""Small driver for ncu: the sampler kernels at BASELINE size (64 x 224^2 x 3)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from dynamic_multiview_3d_b200 import _lib  # noqa: E402

B, H, C = 64, 224, 3
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
data = torch.rand((B, H, H, C), device=dev, generator=g)
flow = (torch.rand((B, H, H, 2), device=dev, generator=g) - 0.5) * 6
go = torch.randn((B, H, H, C), device=dev, generator=g)
out = torch.empty_like(data)
gw = torch.empty_like(flow)
gd = torch.empty_like(data)
ws = torch.empty(_lib.load().dmv_sampler_bwd_workspace_size(B, H, H, C, H, H), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _lib.call("dmv_sampler_fwd", data.data_ptr(), flow.data_ptr(), out.data_ptr(), None, None, B, H, H, C, H, H, 1, st)
    _lib.call("dmv_sampler_bwd", data.data_ptr(), flow.data_ptr(), go.data_ptr(), gd.data_ptr(), gw.data_ptr(), B, H, H, C, H, H, 1,
              ws.data_ptr(), ws.numel(), st)
torch.cuda.synchronize()
print("ok", float(out.sum()), float(gw.abs().sum()), float(gd.abs().sum()))
