#!/bin/bash
# Multi-GPU A/B sweep (run with `gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_dp_sweep.sh 2'`, then 8): eager launches
# vs the step graph with the NCCL exchanges captured (the default), gradient bucket size, how many CTAs NCCL may take, and the
# SM budget of the persistent kernels (the all-reduce moves 392 MB per ~17 ms step: it needs little bandwidth, its CTAs cost SMs).
N=${1:-2}
mkdir -p gpurun_out
run() {   # name, extra env (as VAR=val ...), extra bench args
  local name=$1; shift
  local envs=$1; shift
  env $envs timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 5 --cpu-baseline 0 --eager-baseline 0 "$@" > gpurun_out/dp${N}_$name.json 2> gpurun_out/dp${N}_$name.err
  echo "dp$N $name rc=$? $(python -c "import json; d=json.load(open('gpurun_out/dp${N}_$name.json')); print(d['ms_per_step'], 'ms', d['value'], 'samples/s e2e', d['e2e']['value'], d['config']['launch_mode'][:24])" 2>/dev/null)"
}
run graph "MMAE_NOP=1"
run eager "MMAE_NOP=1" --graph 0
run graph_sm132 "MMAE_NOP=1" --sm-budget 132
run graph_sm140 "MMAE_NOP=1" --sm-budget 140
run graph_ctas8 "NCCL_MAX_CTAS=8"
run graph_ctas8_sm140 "NCCL_MAX_CTAS=8" --sm-budget 140
run graph_bucket16 "MMAE_BUCKET_MB=16"
run graph_bucket128 "MMAE_BUCKET_MB=128"
