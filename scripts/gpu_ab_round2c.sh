#!/bin/bash
# A/B visit: GPU suite, then short bench lines (no CPU / eager baselines) under the switches being compared.
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x --durations=3 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
B="python bench.py --steps 20 --warmup 5 --cpu-baseline 0 --eager-baseline 0"
run() { name=$1; shift; env "$@" timeout 120 $B > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; echo "$name rc=$? $(python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_$name.json"))
    print("ms/step %.3f  e2e %.1f  gemm TF/s %.1f (frac %.3f) launches/step %s" % (d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["gpu_launches"] // 20))
except Exception as e:
    print("no line:", e)
PY
)"; }
run default MMAE_X=0
run no_chain MMAE_BLOCK_CHAIN=0
run no_chain_no_shared MMAE_BLOCK_CHAIN=0 MMAE_SHARED_CTX=0
run default_again MMAE_X=0
for c in 1 2 4 8; do echo "depth std min cluster $c: $(MMAE_DEPTH_STD_MIN_CLUSTER=$c timeout 60 python scripts/gpu_time_depth_standardize.py 2>&1 | grep standardize | cut -c1-90 | tr '\n' '|')"; done
