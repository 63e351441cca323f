#!/bin/bash
# Which form of the fused FC weight-gradient + Adam kernel the captured step prefers: the persistent streaming kernel
# (variant 9: 194 KB of shared memory per SM, nothing of the dgrad chain co-resides) or the generic one-tile-per-CTA
# kernels (variants 0 / 1: 27 KB, short-lived CTAs the high-priority chain can slot in between).
set -u
mkdir -p gpurun_out
T=r02s3a
for v in 9 1 0; do
  DMV_FC_ADAM_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_v$v.json 2> gpurun_out/${T}_bench_v$v.err; echo "bench v$v exit $?"
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_v$v.json')); print('variant $v', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
DMV_FC_ADAM_VARIANT=1 DMV_FUSE_FC_ADAM=1 DMV_DEFER_ADAM=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_v1_fuseall.json 2> gpurun_out/${T}_bench_v1_fuseall.err; echo "bench v1 fuseall exit $?"
python -c "import json; d=json.load(open('gpurun_out/${T}_bench_v1_fuseall.json')); print('variant 1 fuse all', d['value'], d['ms_per_step'], d['e2e']['value'])"
DMV_FC_ADAM_VARIANT=1 timeout 200 python tools/timeline.py gpurun_out/${T}_timeline_v1.txt > gpurun_out/${T}_tl.log 2>&1; echo "tl exit $?"
