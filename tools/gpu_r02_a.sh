#!/bin/bash
# Round 2, call A: the whole GPU test suite (incl. the 224^2 / batch-64 parity tests), the bench line of every BASELINE
# config at N=1, the reference arm, the launch list of one step.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/r02_smi.txt 2>&1
nproc > gpurun_out/r02_nproc.txt
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r02_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/r02_summary.txt
tail -15 gpurun_out/r02_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/r02_summary.txt
for c in 2 3 4 5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r02_bench_c$c.json 2> gpurun_out/r02_bench_c$c.err; echo "bench c$c exit $?" | tee -a gpurun_out/r02_summary.txt
  cut -c1-700 gpurun_out/r02_bench_c$c.json
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref exit $?" | tee -a gpurun_out/r02_summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-micro --eager > gpurun_out/r02_ncu_launches.log 2>&1
echo "ncu launches exit $?" | tee -a gpurun_out/r02_summary.txt
