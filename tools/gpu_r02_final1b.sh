#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02f
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/${T}_summary.txt
tail -4 gpurun_out/${T}_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/${T}_summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench c2 exit $?" | tee -a gpurun_out/${T}_summary.txt
for c in 3 4 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/${T}_bench_c$c.json 2> gpurun_out/${T}_bench_c$c.err; echo "bench c$c exit $?" | tee -a gpurun_out/${T}_summary.txt
done
for c in 2 3 4 5; do python -c "import json; d=json.load(open('gpurun_out/${T}_bench_c$c.json')); r=d['roofline'] or {}; print($c, d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], r.get('kernel'), r.get('frac'), r.get('params'))"; done
python -c "
import torch
print('max mem', torch.cuda.max_memory_allocated())"
