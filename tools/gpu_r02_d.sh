#!/bin/bash
# call D (N GPUs): exchange kernel v2 -- modes agree, raw timings, bench sharded vs fused
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29611 tools/check_dp.py > gpurun_out/r02d_check_dp_$N.log 2>&1; echo "check_dp exit $?" | tee gpurun_out/r02d_summary_$N.txt
tail -4 gpurun_out/r02d_check_dp_$N.log
for ct in 0 96 32; do
DMV_DP_CTAS=$ct CHUNK_MB=128 timeout 600 $TR --master-port 29612 tools/time_comm.py > gpurun_out/r02d_time_comm_${N}_ctas$ct.log 2>&1; echo "time_comm exit $?" | tee -a gpurun_out/r02d_summary_$N.txt
grep "world" gpurun_out/r02d_time_comm_${N}_ctas$ct.log
done
for mode in sharded fused; do
  for ct in 0 64; do
    if [[ $mode == sharded && $ct != 0 ]]; then continue; fi
    DMV_DP_CTAS=$ct DMV_DP_MODE=$mode DMV_DP_MULTICAST=0 timeout 600 $TR --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 --no-micro --no-cpu-baseline \
      > gpurun_out/r02d_bench_${N}gpu_${mode}_ctas$ct.json 2> gpurun_out/r02d_bench_${N}gpu_${mode}_ctas$ct.err
    echo "bench $mode ctas$ct exit $?" | tee -a gpurun_out/r02d_summary_$N.txt
    python -c "import json,sys; d=json.load(open('gpurun_out/r02d_bench_${N}gpu_${mode}_ctas$ct.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['exchange'])"
  done
done
