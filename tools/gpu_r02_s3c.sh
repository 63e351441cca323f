#!/bin/bash
# Co-residency of the input-gradient chain with the persistent fused FC update: ring depth of the streaming kernel
# (DMV_FC_ADAM_STAGES: 3 = 194 KB, 2 = 138 KB per SM) x shared-memory budget of the two-per-SM igemm form
# (DMV_IGEMM_SMEM2_KB), plus the tail lane (DMV_TAIL_LANE: e0 / small-FC weight gradients off lane 0).
set -u
mkdir -p gpurun_out
T=r02s3c
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_$tag.json 2> gpurun_out/${T}_bench_$tag.err; echo "bench $tag exit $?"
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_$tag.json')); print('$tag', d['value'], d['ms_per_step'], d['e2e']['value'], d['final_loss'])"
}
DMV_FC_ADAM_STAGES=2 timeout 300 python -m pytest tests/test_layers_gpu.py -m gpu -q -x -k "adam or act_bwd or bias" --tb=short 2>&1 | tail -3
timeout 300 python -m pytest tests/test_layers_gpu.py -m gpu -q -x -k "adam or act_bwd or bias" --tb=short 2>&1 | tail -3
run default x=1
run tail0 DMV_TAIL_LANE=0
run st2_ig84 DMV_FC_ADAM_STAGES=2 DMV_IGEMM_SMEM2_KB=84
run st2_ig106 DMV_FC_ADAM_STAGES=2
run st3_ig84 DMV_IGEMM_SMEM2_KB=84
run st2_ig60 DMV_FC_ADAM_STAGES=2 DMV_IGEMM_SMEM2_KB=60
DMV_FC_ADAM_STAGES=2 DMV_IGEMM_SMEM2_KB=84 timeout 200 python tools/timeline.py gpurun_out/${T}_timeline_st2_ig84.txt > gpurun_out/${T}_tl.log 2>&1; echo "tl exit $?"
timeout 200 python tools/timeline.py gpurun_out/${T}_timeline_default.txt > gpurun_out/${T}_tl2.log 2>&1; echo "tl2 exit $?"
