#!/bin/bash
# Round-trip on a B200 box: parity tests (one process per file so a fault in one kernel does not
# poison the rest), the bench line, the reference arm, and the ncu launch list / sampler capture.
# Usage (from the repo root, under gpurun):  bash tools/gpu_suite.sh [tests|bench|ncu|all]
set -u
mkdir -p gpurun_out
what=${1:-all}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
if [[ $what == all || $what == tests ]]; then
  for f in test_sampler_gpu test_layers_gpu test_tc_gpu test_model_gpu; do
    timeout 900 python -m pytest tests/$f.py -m gpu -q -x --tb=short > gpurun_out/$f.log 2>&1
    echo "$f exit $?" | tee -a gpurun_out/summary.txt
    tail -5 gpurun_out/$f.log
  done
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt
fi
if [[ $what == all || $what == bench ]]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt
  tail -c 3000 gpurun_out/bench.json
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref exit $?" | tee -a gpurun_out/summary.txt
  tail -c 600 gpurun_out/bench_ref.json
fi
if [[ $what == all || $what == ncu ]]; then
  timeout 300 python tools/prof_sampler.py > gpurun_out/prof_plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:sampler -c 6 -f -o gpurun_out/sampler_prof python tools/prof_sampler.py > gpurun_out/ncu_sampler.log 2>&1
  echo "ncu sampler exit $?" | tee -a gpurun_out/summary.txt
  timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-micro > gpurun_out/bench_short.json 2>&1 &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-micro --eager > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches exit $?" | tee -a gpurun_out/summary.txt
  # dominant kernel of the step (Adam) + the tensor-core kernels, one --set full capture each
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:adam_multi -c 1 -f -o gpurun_out/adam_prof env DMV_OVERLAP_ADAM=0 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-micro --eager > gpurun_out/ncu_adam.log 2>&1
  echo "ncu adam exit $?" | tee -a gpurun_out/summary.txt
  timeout 300 python tools/prof_conv.py > gpurun_out/prof_conv_plain.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"halo|wgrad|igemm" -c 8 -f -o gpurun_out/conv_prof python tools/prof_conv.py > gpurun_out/ncu_conv.log 2>&1
  echo "ncu conv exit $?" | tee -a gpurun_out/summary.txt
fi
