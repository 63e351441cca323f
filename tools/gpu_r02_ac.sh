#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02ac
for cfg in "1 0" "auto auto" "fc1/Matrix,a3/Matrix,a4/Matrix auto" "fc1/Matrix,a5/Matrix auto"; do
  set -- $cfg
  tag=fuse$(echo $1 | tr -d '/,' | sed 's/Matrix//g')_defer$2
  DMV_FUSE_FC_ADAM=$1 DMV_DEFER_ADAM=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_$tag.json 2> gpurun_out/${T}_bench_$tag.err; echo "bench $tag exit $?" | tee -a gpurun_out/${T}_summary.txt
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['final_loss'])"
done
