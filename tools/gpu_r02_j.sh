#!/bin/bash
# call J (8 GPUs): FC-gradient lane and exchange grid A/B in situ, modes agree
set -u
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {
  name=$1; shift
  env "$@" timeout 500 $TR --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 --no-micro --no-cpu-baseline > gpurun_out/r02j_bench_${N}gpu_$name.json 2> gpurun_out/r02j_bench_${N}gpu_$name.err
  echo "bench $name exit $?" | tee -a gpurun_out/r02j_summary_$N.txt
  python -c "import json,sys; d=json.load(open('gpurun_out/r02j_bench_${N}gpu_$name.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['exchange'])"
}
run lane1_ctas0 DMV_FC_LANE=1 DMV_DP_CTAS=0
run lane0_ctas0 DMV_FC_LANE=0 DMV_DP_CTAS=0
run lane1_ctas148 DMV_FC_LANE=1 DMV_DP_CTAS=148
run lane1_ctas64 DMV_FC_LANE=1 DMV_DP_CTAS=64
timeout 400 $TR --master-port 29611 tools/check_dp.py > gpurun_out/r02j_check_dp_$N.log 2>&1; echo "check_dp exit $?" | tee -a gpurun_out/r02j_summary_$N.txt
grep -E "check_dp ok|run-to-run|vs allreduce" gpurun_out/r02j_check_dp_$N.log | sort -u | head -8
