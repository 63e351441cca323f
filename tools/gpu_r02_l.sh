#!/bin/bash
# Full validation of the current tree on one B200: every -m gpu test, smoke, the bench line, a fresh launch list.
set -u
mkdir -p gpurun_out
T=r02l
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/${T}_summary.txt
tail -8 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/${T}_summary.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err; echo "bench exit $?" | tee -a gpurun_out/${T}_summary.txt
python -c "import json; d=json.load(open('gpurun_out/${T}_bench_c2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-micro --eager > gpurun_out/${T}_ncu_launches.log 2>&1
echo "ncu launches exit $?" | tee -a gpurun_out/${T}_summary.txt
