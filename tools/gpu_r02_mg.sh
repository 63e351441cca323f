#!/bin/bash
# Multi-GPU artefacts of the final build.  usage: gpu_r02_mg.sh N "configs" [check]
set -u
N=$1; CFGS=$2; CHECK=${3:-}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -n "$CHECK" ]; then
  timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q --tb=short > gpurun_out/r02mg_pytest_dp_$N.log 2>&1; echo "pytest dp exit $?" | tee -a gpurun_out/r02mg_summary_$N.txt
  tail -3 gpurun_out/r02mg_pytest_dp_$N.log
fi
for c in $CFGS; do
  timeout 500 $TR --master-port 29613 bench.py --gpus $N --config $c --steps 20 --warmup 5 --no-micro --no-cpu-baseline > gpurun_out/r02mg_bench_${N}gpu_c$c.json 2> gpurun_out/r02mg_bench_${N}gpu_c$c.err
  echo "bench N=$N c=$c exit $?" | tee -a gpurun_out/r02mg_summary_$N.txt
  python -c "import json,sys; d=json.load(open('gpurun_out/r02mg_bench_${N}gpu_c$c.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['exchange'])"
done
