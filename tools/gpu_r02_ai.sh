#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02ai
for cfg in "32 0" "64 0" "16 0" "32 120" "32 96" "32 74"; do
  set -- $cfg
  tag=chunk$1_fcctas$2
  DMV_DP_CHUNK_MB=$1 DMV_FC_ADAM_CTAS=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_$tag.json 2> gpurun_out/${T}_bench_$tag.err; echo "bench $tag exit $?"
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_$tag.json')); print('$tag', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
