#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02v
for v in 0 1; do
  if [ $v = 1 ]; then export DMV_NO_THIN_MMA=1; fi
  echo "DMV_NO_THIN_MMA=$v" | tee -a gpurun_out/${T}_time_thin.txt
  timeout 300 python tools/time_layers.py "e0 c5s2" 2>&1 | tee -a gpurun_out/${T}_time_thin.txt
  timeout 300 python tools/time_layers.py "flow" 2>&1 | tee -a gpurun_out/${T}_time_thin.txt
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_nothin$v.json 2> gpurun_out/${T}_bench_nothin$v.err; echo "bench nothin$v exit $?" | tee -a gpurun_out/${T}_summary.txt
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_nothin$v.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['final_loss'])"
done
