#!/bin/bash
set -u
mkdir -p gpurun_out
for c in 5 4; do
  DMV_FUSE_FC_ADAM=0 timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02ad_bench_c${c}_fuse0.json 2> gpurun_out/r02ad_bench_c${c}_fuse0.err
  python -c "import json; d=json.load(open('gpurun_out/r02ad_bench_c${c}_fuse0.json')); print($c, 'fuse0', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
