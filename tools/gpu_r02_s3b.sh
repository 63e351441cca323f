#!/bin/bash
# Stream priority levels (chain > conv weight-gradient lane > FC lane / Adam) x form of the fused FC kernel.
set -u
mkdir -p gpurun_out
T=r02s3b
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_$tag.json 2> gpurun_out/${T}_bench_$tag.err; echo "bench $tag exit $?"
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_$tag.json')); print('$tag', d['value'], d['ms_per_step'], d['e2e']['value'], d['final_loss'])"
}
run v9_prio DMV_FC_ADAM_VARIANT=9
run v1_prio DMV_FC_ADAM_VARIANT=1
run v1_lane0same DMV_FC_ADAM_VARIANT=1 DMV_LANE0_PRIORITY=0
run v0_prio DMV_FC_ADAM_VARIANT=0
run v1_prio_fuseall DMV_FC_ADAM_VARIANT=1 DMV_FUSE_FC_ADAM=1 DMV_DEFER_ADAM=0
run v1_prio_fusefc1 DMV_FC_ADAM_VARIANT=1 DMV_FUSE_FC_ADAM=fc1/Matrix
DMV_FC_ADAM_VARIANT=1 timeout 200 python tools/timeline.py gpurun_out/${T}_timeline_v1_prio.txt > gpurun_out/${T}_tl.log 2>&1; echo "tl exit $?"
