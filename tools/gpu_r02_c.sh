#!/bin/bash
# Round 2, call C (N GPUs, default 2): data-parallel modes agree; raw exchange timings; bench at N in each mode.
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r02c_topo_$N.txt 2>&1
timeout 600 $TR --master-port 29611 tools/check_dp.py > gpurun_out/r02c_check_dp_$N.log 2>&1; echo "check_dp exit $?" | tee gpurun_out/r02c_summary_$N.txt
tail -12 gpurun_out/r02c_check_dp_$N.log
CHUNK_MB=128 timeout 600 $TR --master-port 29612 tools/time_comm.py > gpurun_out/r02c_time_comm_$N.log 2>&1; echo "time_comm exit $?" | tee -a gpurun_out/r02c_summary_$N.txt
grep "world" gpurun_out/r02c_time_comm_$N.log
for mode in sharded fused; do
  for mc in 0 1; do
    if [[ $mode == sharded && $mc == 1 ]]; then continue; fi
    DMV_DP_MODE=$mode DMV_DP_MULTICAST=$mc timeout 600 $TR --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 --no-micro --no-cpu-baseline \
      > gpurun_out/r02c_bench_${N}gpu_${mode}_mc$mc.json 2> gpurun_out/r02c_bench_${N}gpu_${mode}_mc$mc.err
    echo "bench $mode mc$mc exit $?" | tee -a gpurun_out/r02c_summary_$N.txt
    python -c "import json,sys; d=json.load(open('gpurun_out/r02c_bench_${N}gpu_${mode}_mc$mc.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['exchange'])"
  done
done
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL DMV_DP_MODE=sharded timeout 600 $TR --master-port 29614 bench.py --gpus $N --steps 3 --warmup 3 --no-micro --no-cpu-baseline \
  > /dev/null 2> gpurun_out/r02c_nccl_debug_$N.err
grep -iE "NVLS|algo|Channel|Connected" gpurun_out/r02c_nccl_debug_$N.err | head -40 > gpurun_out/r02c_nccl_algo_$N.txt
wc -l gpurun_out/r02c_nccl_algo_$N.txt
