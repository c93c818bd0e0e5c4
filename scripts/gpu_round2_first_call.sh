#!/bin/bash
# First GPU-box visit of the next round: everything that was prepared after round 1's GPU budget ran out, in one call.
#   1. GPU suite + smoke (regression check of the committed state)
#   2. A/B of MMAE_GEMM_BALANCED_GRID (persistent GEMM grid = ceil(items / rounds)) on the headline bench
#   3. K-sweep diagnostic of the K = 256 decoder GEMM shapes
#   4. experimental depth-standardisation variant 2 against the oracle and variant 1 (correctness + timing)
#   5. experimental persistent encoder-attention forward: bit-identity with the default kernel + timing (last: may hang)
# Logs land in gpurun_out/; copy what is kept into profiles/rNN_*.
#     gpurun --timeout 600 -- 'bash scripts/gpu_round2_first_call.sh'
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
for v in 0 1; do
  MMAE_GEMM_BALANCED_GRID=$v timeout 90 python bench.py --steps 20 --warmup 5 --cpu-baseline 0 \
      --gemm-shapes gpurun_out/gemm_shapes_balanced$v.txt > gpurun_out/bench_balanced$v.json 2> gpurun_out/bench_balanced$v.err
  echo "bench balanced_grid=$v rc=$?"; cut -c1-200 gpurun_out/bench_balanced$v.json
done
timeout 120 python scripts/gpu_diag_gemm_smallk.py > gpurun_out/gemm_smallk_diag.log 2>&1; echo "small-K diag rc=$?"; tail -40 gpurun_out/gemm_smallk_diag.log
# the two unproven kernels go last (cluster / mbarrier protocol bugs show up as hangs: bounded by `timeout`)
timeout 90 python scripts/gpu_check_depth_standardize_v2.py > gpurun_out/depth_standardize_v2.log 2>&1; echo "depth v2 rc=$?"; tail -6 gpurun_out/depth_standardize_v2.log
# a barrier-protocol bug in an unproven tcgen05 kernel shows up as a hang: bounded by `timeout`, nothing after it
timeout 120 python scripts/gpu_check_attention_v2.py > gpurun_out/attention_v2.log 2>&1; echo "attention v2 rc=$? (124 = hang)"; tail -10 gpurun_out/attention_v2.log
