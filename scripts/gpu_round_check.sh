#!/bin/bash
# One GPU-box visit: the whole GPU test suite, smoke(), the headline bench line, the cfg-4 / cfg-5 workloads, the step
# with truncated depth standardisation, and the depth-standardisation kernel timing.  Logs land in gpurun_out/.
mkdir -p gpurun_out
timeout 150 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
timeout 40 python scripts/gpu_time_depth_standardize.py > gpurun_out/depth_standardize_timing.log 2>&1; echo "depth timing rc=$?"
timeout 200 python bench.py --gemm-shapes gpurun_out/gemm_shapes.txt > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "bench rc=$?"
timeout 60 python bench.py --workload cfg4 --steps 10 --cpu-baseline 0 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "bench cfg4 rc=$?"
timeout 60 python bench.py --workload cfg5 --steps 10 --cpu-baseline 0 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "bench cfg5 rc=$?"
timeout 60 python bench.py --standardize-depth 1 --steps 10 --cpu-baseline 0 > gpurun_out/bench_cfg2_depthstd.json 2> gpurun_out/bench_cfg2_depthstd.err; echo "bench depthstd rc=$?"
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/depth_standardize_timing.log; for f in gpurun_out/bench_*.json; do echo $f; cut -c1-330 $f; done
