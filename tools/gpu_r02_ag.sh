#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_dp_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/r02ag_bench_c2.json 2> gpurun_out/r02ag_bench_c2.err; echo "bench exit $?"
python -c "import json; d=json.load(open('gpurun_out/r02ag_bench_c2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['final_loss'])"
