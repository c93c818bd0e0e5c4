#!/bin/bash
# One GPU-box visit that produces the evidence kept under profiles/ (every step bounded by a timeout):
#   1. the whole GPU suite + smoke()            2. the headline bench line (with the eager-GPU and CPU baselines)
#   3. launch list of one eager step (ncu)      4. `ncu --set full` of the hot kernels (scripts/gpu_ncu_targets.py)
#   5. attention kernel timings                 6. cfg-4 / cfg-5 bench lines
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py --gemm-shapes gpurun_out/gemm_shapes.txt > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_cfg2.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --graph 0 --cpu-baseline 0 --eager-baseline 0 > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_targets \
    python scripts/gpu_ncu_targets.py > gpurun_out/ncu_targets.log 2>&1; echo "full capture rc=$?"
timeout 120 python scripts/gpu_check_attention_ws.py all > gpurun_out/attn_ws_all.log 2>&1; echo "attention rc=$?"; grep "^time" gpurun_out/attn_ws_all.log
timeout 90 python bench.py --workload cfg4 --steps 10 --cpu-baseline 0 --eager-baseline 0 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 rc=$?"; cut -c1-160 gpurun_out/bench_cfg4.json
timeout 90 python bench.py --workload cfg5 --steps 10 --cpu-baseline 0 --eager-baseline 0 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 rc=$?"; cut -c1-160 gpurun_out/bench_cfg5.json
