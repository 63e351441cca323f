#!/bin/bash
# call E (1 GPU): full GPU suite after the transposed-lane sampler, bench line, ncu capture of the sampler kernels + the fused train kernel
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02e_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/r02e_summary.txt
tail -12 gpurun_out/r02e_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02e_bench_c2.json 2> gpurun_out/r02e_bench_c2.err; echo "bench exit $?" | tee -a gpurun_out/r02e_summary.txt
python -c "import json; d=json.load(open('gpurun_out/r02e_bench_c2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['sampler'])"
timeout 300 python tools/prof_sampler.py > gpurun_out/r02e_prof_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sampler -c 6 -f -o gpurun_out/r02e_sampler_prof python tools/prof_sampler.py > gpurun_out/r02e_ncu_sampler.log 2>&1
echo "ncu sampler exit $?" | tee -a gpurun_out/r02e_summary.txt
timeout 300 python tools/time_warp_loss.py > gpurun_out/r02e_time_warp_loss.log 2>&1; tail -5 gpurun_out/r02e_time_warp_loss.log
