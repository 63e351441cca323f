#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02s
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/${T}_summary.txt
tail -6 gpurun_out/${T}_pytest.log
for i in 1 2; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_c2_$i.json 2> gpurun_out/${T}_bench_c2_$i.err; echo "bench c2 $i exit $?" | tee -a gpurun_out/${T}_summary.txt
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_c2_$i.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['final_loss'])"
done
for c in 3 4 5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_c$c.json 2> gpurun_out/${T}_bench_c$c.err; echo "bench c$c exit $?" | tee -a gpurun_out/${T}_summary.txt
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_c$c.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['final_loss'])"
done
