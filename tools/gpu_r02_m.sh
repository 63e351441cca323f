#!/bin/bash
# Fused FC weight-gradient + Adam kernel: tests, then A/B bench lines on one box.
set -u
mkdir -p gpurun_out
T=r02m
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -k "adam or fused or train_step or checkpoint or graphed" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/${T}_summary.txt
tail -8 gpurun_out/${T}_pytest.log
for cfg in "1 0 0" "1 1 0" "1 1 96" "1 1 64" "1 1 0 G" "0 0 0"; do
  set -- $cfg
  tag=fuse$1_lane$2_ctas$3${4:-}
  DMV_FUSE_FC_ADAM=$1 DMV_FC_LANE=$2 DMV_FC_ADAM_CTAS=$3 DMV_FC_ADAM_GENERIC=$([ "${4:-}" = G ] && echo 1 || echo 0) timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_$tag.json 2> gpurun_out/${T}_bench_$tag.err; echo "bench $tag exit $?" | tee -a gpurun_out/${T}_summary.txt
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['ms'], d['final_loss'])"
done
