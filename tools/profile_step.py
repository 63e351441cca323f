"""Per-call CUDA-event profile of one eager train step at the BASELINE config (B=64, 224^2)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import dynamic_multiview_3d_b200 as pkg  # noqa: E402
from dynamic_multiview_3d_b200 import functional as F  # noqa: E402
from dynamic_multiview_3d_b200.synthetic import make_batch  # noqa: E402

B = int(os.environ.get("B", 64))
model = pkg.AppearanceFlowModel({"batch_size": B, "learning_rate": 1e-4, "image_size": 224, "viewpoint_dim": 19})
b = make_batch(B, 224, "onehot19")
args = [torch.from_numpy(b[k]).cuda() for k in ("image0", "image1", "disp")]
for _ in range(3):
    model.train_step(*args)
with F.profile_calls() as prof:
    model.train_step(*args)
agg = {}
for name, tag, ms in prof.records:
    k = (name, tag)
    agg[k] = agg.get(k, 0.0) + ms
tot = sum(agg.values())
rows = sorted(agg.items(), key=lambda kv: -kv[1])
print("total of timed calls: %.3f ms over %d calls" % (tot, len(prof.records)))
for (name, tag), ms in rows:
    print("%-26s %-16s %8.1f us  %5.1f%%" % (name, tag, ms * 1e3, 100 * ms / tot))
by = {}
for (name, tag), ms in agg.items():
    by[name] = by.get(name, 0.0) + ms
print("---- by entry point")
for name, ms in sorted(by.items(), key=lambda kv: -kv[1]):
    print("%-26s %8.1f us  %5.1f%%" % (name, ms * 1e3, 100 * ms / tot))
if len(sys.argv) > 1:
    json.dump({"total_ms": tot, "calls": [[n, t, ms] for (n, t), ms in rows]}, open(sys.argv[1], "w"), indent=0)
