#!/bin/bash
# A/B visit: CTA-pair GEMM kernels on more shapes (MMAE_GEMM_PAIR_NK = smallest N*K that may use them), per-shape GEMM tables
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --cpu-baseline 0 --eager-baseline 0"
run() { name=$1; shift; env "$@" timeout 120 $B --gemm-shapes gpurun_out/shapes_$name.txt > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err; echo "$name rc=$? $(python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_$name.json"))
    print("ms/step %.3f  e2e %.1f  gemm TF/s %.1f (frac %.3f) enc %.1f launches/step %s" % (d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["encoder_tc"]["encoder_gemm_tflops"], d["gpu_launches"] // 20))
except Exception as e:
    print("no line:", e)
PY
)"; }
run default MMAE_X=0
run pair_nk_1p7m MMAE_GEMM_PAIR_NK=1700000
run pair_nk_500k MMAE_GEMM_PAIR_NK=500000
run default_again MMAE_X=0
timeout 100 python bench.py --steps 10 --cpu-baseline 0 --eager-baseline 0 --standardize-depth 1 > gpurun_out/bench_cfg2_depthstd.json 2> gpurun_out/bench_cfg2_depthstd.err; echo "depthstd rc=$? $(cut -c1-140 gpurun_out/bench_cfg2_depthstd.json)"
