#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02y
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/${T}_summary.txt
tail -6 gpurun_out/${T}_pytest.log | cut -c1-300
for v in 1 0; do
  DMV_PREPACK=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_prepack$v.json 2> gpurun_out/${T}_bench_prepack$v.err; echo "bench prepack$v exit $?" | tee -a gpurun_out/${T}_summary.txt
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_prepack$v.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['final_loss'])"
done
timeout 600 python tools/timeline.py gpurun_out/${T}_timeline.txt > gpurun_out/${T}_tl.log 2>&1; echo "tl exit $?"
