"""The sampler microbenchmark of bench.py on its own (64 x 224^2 x 3, jittered and smooth flows): forward, grad wrt flow,
grad wrt flow + source.  Usage (GPU box): python tools/time_sampler.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    torch.cuda.set_device(0)
    print(json.dumps(bench.sampler_microbench(torch, bench.peaks()), indent=1))
