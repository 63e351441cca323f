#!/bin/bash
# Final single-GPU artefacts of round 2: every -m gpu test, smoke, bench lines of configs 2-5, the reference arm, the ncu
# launch list of an eager step, one --set full capture of the dominant kernels.
set -u
mkdir -p gpurun_out
T=r02f
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/${T}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/${T}_summary.txt
tail -4 gpurun_out/${T}_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/${T}_summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench c2 exit $?" | tee -a gpurun_out/${T}_summary.txt
for c in 3 4 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/${T}_bench_c$c.json 2> gpurun_out/${T}_bench_c$c.err; echo "bench c$c exit $?" | tee -a gpurun_out/${T}_summary.txt
done
for c in 2 3 4 5; do python -c "import json; d=json.load(open('gpurun_out/${T}_bench_c$c.json')); print($c, d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['roofline'].get('kernel'), d['roofline'].get('frac'))"; done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref exit $?" | tee -a gpurun_out/${T}_summary.txt
timeout 600 python tools/time_layers.py > gpurun_out/${T}_time_layers.txt 2>&1; echo "time_layers exit $?" | tee -a gpurun_out/${T}_summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-micro --eager > gpurun_out/${T}_ncu_launches.log 2>&1
echo "ncu launches exit $?" | tee -a gpurun_out/${T}_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:adam_multi -c 1 -f -o gpurun_out/${T}_adam env DMV_OVERLAP_ADAM=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-micro --eager > gpurun_out/${T}_ncu_adam.log 2>&1
echo "ncu adam exit $?" | tee -a gpurun_out/${T}_summary.txt
