#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02f_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/r02f_summary.txt
tail -12 gpurun_out/r02f_pytest.log
