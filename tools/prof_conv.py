"""Driver for ncu: the biggest tensor-core conv layer (e0_0 / d1_0: 5x5, 32->32 at 112^2, B=64), fwd + dgrad + wgrad."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from dynamic_multiview_3d_b200 import functional as F  # noqa: E402
from dynamic_multiview_3d_b200.variables import VariableStore  # noqa: E402

B, H, C = 64, 112, 32
st = VariableStore(torch.device("cuda:0"))
w = st.get("w", (5, 5, C, C), "truncated_normal", 0.05)
bb = st.get("b", (C,), "zeros")
x = torch.randn((B, H, H, C), device="cuda").to(torch.bfloat16).requires_grad_(True)
for _ in range(2):
    y = F.conv2d(x, w, bb, 1, "lrelu", "auto")
    y.backward(torch.randn_like(y))
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
