#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02x
timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -x -k "thin or forward_parity or colordepth or multiobject" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/${T}_summary.txt
tail -4 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 300 python tools/time_layers.py "5s2 224" 2>&1 | tee gpurun_out/${T}_time_thin.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?" | tee -a gpurun_out/${T}_summary.txt
python -c "import json; d=json.load(open('gpurun_out/${T}_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['final_loss'])"
for k in thin_conv thin_deconv thin_wgrad_mma; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 3 -c 1 -f -o gpurun_out/${T}_$k python tools/time_layers.py "5s2 224" > gpurun_out/${T}_ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
