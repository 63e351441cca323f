#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02s3g
timeout 600 python -m pytest tests/test_sampler_gpu.py -m gpu -q -x --tb=short > gpurun_out/${T}_pytest_sampler.log 2>&1; echo "pytest sampler exit $?"
tail -2 gpurun_out/${T}_pytest_sampler.log | cut -c1-300
timeout 200 python tools/time_sampler.py > gpurun_out/${T}_time_sampler.txt 2>&1; echo "time exit $?"
grep -A2 bwd_both gpurun_out/${T}_time_sampler.txt | head -30
timeout 200 python tools/time_sampler_variants.py jitter > gpurun_out/${T}_variants_jitter.txt 2>&1; echo "variants exit $?"
timeout 200 python tools/time_sampler_variants.py smooth > gpurun_out/${T}_variants_smooth.txt 2>&1; echo "variants exit $?"
cat gpurun_out/${T}_variants_jitter.txt gpurun_out/${T}_variants_smooth.txt
timeout 200 python tools/time_warp_loss.py > gpurun_out/${T}_time_warp_loss.txt 2>&1; echo "warp_loss exit $?"
tail -6 gpurun_out/${T}_time_warp_loss.txt
