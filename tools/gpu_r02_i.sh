#!/bin/bash
# call I (8 GPUs): raw exchange timings, the bench of every BASELINE config in the fused mode, config 2 in the NCCL mode, modes agree
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
CHUNK_MB=128 timeout 400 $TR --master-port 29612 tools/time_comm.py > gpurun_out/r02i_time_comm_$N.log 2>&1; echo "time_comm exit $?" | tee gpurun_out/r02i_summary_$N.txt
grep "world" gpurun_out/r02i_time_comm_$N.log
run() {  # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 500 $TR --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 --no-micro --no-cpu-baseline $BARGS > gpurun_out/r02i_bench_${N}gpu_$name.json 2> gpurun_out/r02i_bench_${N}gpu_$name.err
  echo "bench $name exit $?" | tee -a gpurun_out/r02i_summary_$N.txt
  python -c "import json,sys; d=json.load(open('gpurun_out/r02i_bench_${N}gpu_$name.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['exchange'])"
}
BARGS="--config 2" run c2_fused DMV_DP_MODE=fused DMV_DP_MULTICAST=0
BARGS="--config 2" run c2_fused_mc DMV_DP_MODE=fused DMV_DP_MULTICAST=1
BARGS="--config 2" run c2_sharded DMV_DP_MODE=sharded
BARGS="--config 3" run c3_fused DMV_DP_MODE=fused DMV_DP_MULTICAST=0
BARGS="--config 4" run c4_fused DMV_DP_MODE=fused DMV_DP_MULTICAST=0
BARGS="--config 5" run c5_fused DMV_DP_MODE=fused DMV_DP_MULTICAST=0
timeout 400 $TR --master-port 29611 tools/check_dp.py > gpurun_out/r02i_check_dp_$N.log 2>&1; echo "check_dp exit $?" | tee -a gpurun_out/r02i_summary_$N.txt
tail -4 gpurun_out/r02i_check_dp_$N.log
