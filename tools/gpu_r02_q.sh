#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02q
for cfg in "1 auto 0" "0 auto 0" "1 0 0" "1 auto 592" "1 auto 296" "1 0 592"; do
  set -- $cfg
  tag=prio$1_defer$2_adamctas$3
  DMV_MAIN_PRIORITY=$1 DMV_DEFER_ADAM=$2 DMV_ADAM_CTAS=$3 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_$tag.json 2> gpurun_out/${T}_bench_$tag.err; echo "bench $tag exit $?" | tee -a gpurun_out/${T}_summary.txt
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['final_loss'])"
done
DMV_MAIN_PRIORITY=1 timeout 600 python tools/timeline.py gpurun_out/${T}_timeline_prio.txt > gpurun_out/${T}_tl.log 2>&1; echo "tl exit $?"
