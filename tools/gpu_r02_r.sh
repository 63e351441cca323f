#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02r
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/${T}_summary.txt
tail -6 gpurun_out/${T}_pytest.log
for cfg in "1 0" "0 0" "1 fc1/Matrix"; do
  set -- $cfg
  tag=branch$1_fuse$(echo $2 | tr -d '/,')
  DMV_VIEW_BRANCH=$1 DMV_FUSE_FC_ADAM=$2 DMV_FC_ADAM_VARIANT=9 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_$tag.json 2> gpurun_out/${T}_bench_$tag.err; echo "bench $tag exit $?" | tee -a gpurun_out/${T}_summary.txt
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['final_loss'])"
done
timeout 600 python tools/timeline.py gpurun_out/${T}_timeline.txt > gpurun_out/${T}_tl.log 2>&1; echo "tl exit $?"
