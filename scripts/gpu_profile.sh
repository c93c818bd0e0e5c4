#!/bin/bash
# One GPU-box visit that refreshes the profiling evidence (B200_PROFILING.md recipe; one GPU, never under torchrun):
#   1. launch list of the training step:  gpurun_out/launches.csv   -> python scripts/summarize_launches.py (here, offline)
#   2. `ncu --set full` of the hot kernels at their step shapes: gpurun_out/prof_targets.ncu-rep
#        -> ncu -i ... --page raw --csv | python scripts/ncu_raw_table.py (here, offline)
# Numbers printed by bench.py under ncu are NOT bench values.
#     gpurun --timeout 600 -- 'bash scripts/gpu_profile.sh'
mkdir -p gpurun_out
# eager launches (a replayed CUDA graph is profiled node by node as well, but the eager step keeps kernel names and order
# identical to the per-shape GEMM table); 3 warm-up + 1 timed + profile/e2e steps: skip the first 4 steps' launches
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --graph 0 --cpu-baseline 0 > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$?"
timeout 240 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/prof_targets \
    python scripts/gpu_ncu_targets.py > gpurun_out/ncu_targets.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/launches.csv gpurun_out/prof_targets.ncu-rep
