#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02aa
timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_parity_full_gpu.py -m gpu -q --tb=short -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/${T}_summary.txt
tail -4 gpurun_out/${T}_pytest.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit $?" | tee -a gpurun_out/${T}_summary.txt
python -c "import json; d=json.load(open('gpurun_out/${T}_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['final_loss'])"
timeout 600 python tools/timeline.py gpurun_out/${T}_timeline.txt > gpurun_out/${T}_tl.log 2>&1; echo "tl exit $?"
