#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02h_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/r02h_summary.txt
tail -12 gpurun_out/r02h_pytest.log
for fuse in 1 0; do
DMV_FUSE_DACT=$fuse timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02h_bench_c2_fuse$fuse.json 2> gpurun_out/r02h_bench_c2_fuse$fuse.err; echo "bench fuse$fuse exit $?" | tee -a gpurun_out/r02h_summary.txt
python -c "import json; d=json.load(open('gpurun_out/r02h_bench_c2_fuse$fuse.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['conv'])"
done
timeout 300 python tools/time_layers.py > gpurun_out/r02h_time_layers.txt 2>&1; tail -3 gpurun_out/r02h_time_layers.txt
