#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02t
timeout 600 python -m pytest tests/test_tc_gpu.py -m gpu -q --tb=short -k "thin" > gpurun_out/${T}_pytest_thin.log 2>&1; echo "pytest thin exit $?" | tee gpurun_out/${T}_summary.txt
tail -15 gpurun_out/${T}_pytest_thin.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${T}_summary.txt
tail -6 gpurun_out/${T}_pytest.log | cut -c1-300
for v in 0 1; do
  if [ $v = 1 ]; then export DMV_NO_THIN_MMA=1; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_nothin$v.json 2> gpurun_out/${T}_bench_nothin$v.err; echo "bench nothin$v exit $?" | tee -a gpurun_out/${T}_summary.txt
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_nothin$v.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['final_loss'])"
done
unset DMV_NO_THIN_MMA
timeout 600 python tools/timeline.py gpurun_out/${T}_timeline.txt > gpurun_out/${T}_tl.log 2>&1; echo "tl exit $?"
