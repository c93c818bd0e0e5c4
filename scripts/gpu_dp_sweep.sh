#!/bin/bash
# Multi-GPU A/B sweep for the next round (run with `gpurun --gpus 2 --timeout 600 -- 'bash scripts/gpu_dp_sweep.sh 2'`, then 8):
# eager launches vs the step graph with the NCCL exchanges captured, gradient bucket size, and how many CTAs NCCL may take
# from the persistent GEMMs (the all-reduce moves 392 MB per ~17 ms step: it needs little bandwidth, its CTAs cost SMs).
N=${1:-2}
mkdir -p gpurun_out
run() {   # name, extra env (as VAR=val ...), extra bench args
  local name=$1; shift
  local envs=$1; shift
  env $envs timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/dp${N}_$name.json 2> gpurun_out/dp${N}_$name.err
  echo "dp$N $name rc=$? $(grep -o '"value": [0-9.]*, "unit"' gpurun_out/dp${N}_$name.json | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/dp${N}_$name.json | head -1)"
}
run eager "MMAE_NOP=1"
run graph "MMAE_NOP=1" --graph 2
run bucket16 "MMAE_BUCKET_MB=16"
run bucket96 "MMAE_BUCKET_MB=96"
run ctas8 "NCCL_MAX_CTAS=8"
run ctas16 "NCCL_MAX_CTAS=16"
run ctas8_graph "NCCL_MAX_CTAS=8" --graph 2
