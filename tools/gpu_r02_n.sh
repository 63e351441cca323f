#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02n
for v in 0 1 4 9; do
  DMV_FC_ADAM_VARIANT=$v timeout 300 python -m pytest tests/test_layers_gpu.py -m gpu -q --tb=line -k "wgrad_adam" 2>&1 | tail -3
  DMV_FC_ADAM_VARIANT=$v timeout 300 python tools/time_fc_adam.py 2>&1 | tee -a gpurun_out/${T}_time_fc_adam.txt
done
for cfg in "4 1" "1 1" "9 1" "1 0"; do
  set -- $cfg
  tag=var$1_lane$2
  DMV_FC_ADAM_VARIANT=$1 DMV_FC_LANE=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_$tag.json 2> gpurun_out/${T}_bench_$tag.err; echo "bench $tag exit $?" | tee -a gpurun_out/${T}_summary.txt
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['roofline']['kernel'], 'frac', d['roofline']['frac'], 'ms', d['roofline']['ms'])"
done
