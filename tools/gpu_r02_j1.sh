#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02j1_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/r02j1_summary.txt
tail -5 gpurun_out/r02j1_pytest.log
for lane in 1 0; do
DMV_FC_LANE=$lane timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/r02j1_bench_c2_lane$lane.json 2> gpurun_out/r02j1_bench_c2_lane$lane.err; echo "bench lane$lane exit $?" | tee -a gpurun_out/r02j1_summary.txt
python -c "import json; d=json.load(open('gpurun_out/r02j1_bench_c2_lane$lane.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])"
done
