#!/bin/bash
# Final single-GPU validation of the round-2 tree: every -m gpu test, smoke, bench lines of configs 2-5, the reference arm.
set -u
mkdir -p gpurun_out
T=r02g
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/${T}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/${T}_summary.txt
tail -4 gpurun_out/${T}_pytest.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/${T}_summary.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench c2 exit $?" | tee -a gpurun_out/${T}_summary.txt
for c in 3 4 5; do
  timeout 300 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/${T}_bench_c$c.json 2> gpurun_out/${T}_bench_c$c.err; echo "bench c$c exit $?" | tee -a gpurun_out/${T}_summary.txt
done
for c in 2 3 4 5; do python -c "import json; d=json.load(open('gpurun_out/${T}_bench_c$c.json')); r=d['roofline'] or {}; print($c, d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], r.get('kernel'), r.get('frac'))"; done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref exit $?" | tee -a gpurun_out/${T}_summary.txt
tail -c 400 gpurun_out/${T}_bench_ref.json
