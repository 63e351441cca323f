#!/bin/bash
# Round 2, call B (1 GPU): full GPU test suite after the parity rework, exchange-kernel emulation test.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02b_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/r02b_summary.txt
tail -40 gpurun_out/r02b_pytest.log
