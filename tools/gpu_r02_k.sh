#!/bin/bash
# call K (N GPUs): exchange grid sweep in situ.  usage: gpu_r02_k.sh N "ctas list"
set -u
N=$1; LIST=$2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for ct in $LIST; do
  DMV_DP_CTAS=$ct timeout 500 $TR --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 --no-micro --no-cpu-baseline > gpurun_out/r02k_bench_${N}gpu_ctas$ct.json 2> gpurun_out/r02k_bench_${N}gpu_ctas$ct.err
  echo "bench N=$N ctas=$ct exit $?" | tee -a gpurun_out/r02k_summary_$N.txt
  python -c "import json,sys; d=json.load(open('gpurun_out/r02k_bench_${N}gpu_ctas$ct.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['exchange'])"
done
