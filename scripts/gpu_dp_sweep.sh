#!/bin/bash
# Multi-GPU A/B sweep (run with `gpurun --gpus 2 --timeout 600 -- 'bash scripts/gpu_dp_sweep.sh 2 [names...]'`): eager launches
# vs the step graph with the NCCL exchanges captured (the default), how many CTAs NCCL may take, and the SM budget of the
# persistent kernels (the all-reduce moves 392 MB per ~17 ms step: it needs little bandwidth, its CTAs cost SMs).
# Every run is bounded by a (60 + 15 N) s timeout and the sweep STOPS at the first run that fails or hangs (multi-GPU minutes are
# charged N-fold).
N=${1:-2}; shift
SEL="$*"
mkdir -p gpurun_out
# A fresh box pages torch / CUDA libraries in on first import; with N ranks starting at once that alone took longer than a
# 75 s run limit at N = 8 (round 2: an 8-GPU visit, charged 8-fold, that measured nothing).  Page everything in ONCE first and
# give the runs a limit that scales with N.
timeout 300 python -c "import torch, multimae_b200._lib as L; L.lib(); print('warm:', torch.__version__, torch.cuda.device_count(), 'GPUs')"
LIMIT=$((60 + 15 * N))
run() {   # name, extra env (as VAR=val ...), extra bench args
  local name=$1; shift
  local envs=$1; shift
  if [ -n "$SEL" ] && ! echo " $SEL " | grep -q " $name "; then return; fi
  env $envs timeout $LIMIT python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 5 --cpu-baseline 0 --eager-baseline 0 "$@" > gpurun_out/dp${N}_$name.json 2> gpurun_out/dp${N}_$name.err
  local rc=$?
  echo "dp$N $name rc=$rc $(python -c "import json; d=json.load(open('gpurun_out/dp${N}_$name.json')); print(d['ms_per_step'], 'ms', d['value'], 'samples/s e2e', d['e2e']['value'], d['config']['launch_mode'][:24])" 2>/dev/null)"
  if [ $rc -ne 0 ]; then echo "STOP: $name failed (rc=$rc)"; grep -v "^\s" gpurun_out/dp${N}_$name.err | tail -5; exit 1; fi
}
run graph "MMAE_NOP=1"
run eager "MMAE_NOP=1" --graph 0
run graph_sm140 "MMAE_NOP=1" --sm-budget 140
run graph_sm132 "MMAE_NOP=1" --sm-budget 132
run graph_ctas8 "NCCL_MAX_CTAS=8"
run graph_ctas16_sm132 "NCCL_MAX_CTAS=16" --sm-budget 132
