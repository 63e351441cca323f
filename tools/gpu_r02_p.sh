#!/bin/bash
set -u
mkdir -p gpurun_out
T=r02p
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -k "adam or fused or deferred or train_step or checkpoint or graphed" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" | tee gpurun_out/${T}_summary.txt
tail -8 gpurun_out/${T}_pytest.log
for cfg in "auto 0" "0 0" "a5/Matrix,a4/Matrix 0" "auto fc1/Matrix"; do
  set -- $cfg
  tag=defer$(echo $1 | tr -d '/,' )_fuse$(echo $2 | tr -d '/,')
  DMV_DEFER_ADAM=$1 DMV_FUSE_FC_ADAM=$2 DMV_FC_ADAM_VARIANT=9 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-micro > gpurun_out/${T}_bench_$tag.json 2> gpurun_out/${T}_bench_$tag.err; echo "bench $tag exit $?" | tee -a gpurun_out/${T}_summary.txt
  python -c "import json; d=json.load(open('gpurun_out/${T}_bench_$tag.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['final_loss'])"
done
DMV_FUSE_FC_ADAM=0 timeout 600 python tools/timeline.py gpurun_out/${T}_timeline_deferred.txt > gpurun_out/${T}_tl.log 2>&1; echo "tl exit $?"
DMV_FUSE_FC_ADAM=0 DMV_DEFER_ADAM=0 timeout 600 python tools/timeline.py gpurun_out/${T}_timeline_immediate.txt > gpurun_out/${T}_tl0.log 2>&1; echo "tl0 exit $?"
