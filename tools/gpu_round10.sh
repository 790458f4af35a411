#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_graph.py -m gpu -q > gpurun_out/test_gemm.log 2>&1; echo "test_gemm+graph rc=$?"; tail -n 4 gpurun_out/test_gemm.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_modules.py -m gpu -q > gpurun_out/test_modules.log 2>&1; echo "test_modules rc=$?"; tail -n 3 gpurun_out/test_modules.log
timeout 300 python tools/gemm_table.py > gpurun_out/gemm_table.log 2>&1; echo "gemm_table rc=$?"; cut -c1-200 gpurun_out/gemm_table.log
timeout 300 python tools/profile_step.py torchprof > gpurun_out/torchprof.log 2>&1; echo "torchprof rc=$?"; sed -n 3,30p gpurun_out/torchprof.log | cut -c1-150
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench.log
