#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02s3d
timeout 600 python -m pytest tests/test_sampler_gpu.py -m gpu -q -x --tb=short > gpurun_out/${T}_pytest_sampler.log 2>&1; echo "pytest sampler exit $?"
tail -5 gpurun_out/${T}_pytest_sampler.log | cut -c1-300
timeout 200 python tools/time_sampler.py > gpurun_out/${T}_time_sampler.txt 2>&1; echo "time exit $?"
grep -A5 bwd_both gpurun_out/${T}_time_sampler.txt | head -30
