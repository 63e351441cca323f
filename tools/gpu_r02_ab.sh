#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02ab
timeout 300 python -m pytest tests/test_layers_gpu.py tests/test_model_gpu.py -m gpu -q --tb=line -k "wgrad_adam or fused_fc" 2>&1 | tail -3
timeout 300 python tools/time_fc_adam.py 2>&1 | tee gpurun_out/${T}_time_fc_adam.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?"
python -c "import json; d=json.load(open('gpurun_out/${T}_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['final_loss'], d['roofline'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fc_wgrad_adam" --launch-skip 3 -c 1 -f -o gpurun_out/${T}_fcadam python tools/time_fc_adam.py > gpurun_out/${T}_ncu_fcadam.log 2>&1; echo "ncu fcadam exit $?"
