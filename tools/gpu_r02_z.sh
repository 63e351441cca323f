#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02z
for cfg in "auto 0" "auto 1" "a5/Matrix,a4/Matrix 0" "a5/Matrix 0"; do
  set -- $cfg
  tag=defer$(echo $1 | tr -d '/,' | sed 's/Matrix//g')_prepack$2
  DMV_DEFER_ADAM=$1 DMV_PREPACK=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_$tag.json 2> gpurun_out/${T}_bench_$tag.err; echo "bench $tag exit $?" | tee -a gpurun_out/${T}_summary.txt
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['final_loss'])"
done
DMV_PREPACK=0 timeout 600 python tools/timeline.py gpurun_out/${T}_timeline.txt > gpurun_out/${T}_tl.log 2>&1; echo "tl exit $?"
