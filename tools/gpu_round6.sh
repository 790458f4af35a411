#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/test_gemm.log 2>&1; echo "test_gemm rc=$?"; tail -n 3 gpurun_out/test_gemm.log
timeout 600 python tools/graph_debug.py > gpurun_out/graph_debug.log 2>&1; echo "graph_debug rc=$?"; grep -v Warning gpurun_out/graph_debug.log | head -50
timeout 900 python -m pytest tests/test_gpu_modules.py -m gpu -q > gpurun_out/test_modules.log 2>&1; echo "test_modules rc=$?"; tail -n 3 gpurun_out/test_modules.log
timeout 900 python tools/gemm_table.py > gpurun_out/gemm_table.log 2>&1; echo "gemm_table rc=$?"; cat gpurun_out/gemm_table.log | cut -c1-250
timeout 900 python bench.py --steps 10 --warmup 3 --no-graph > gpurun_out/bench_nograph.log 2>&1; echo "bench nograph rc=$?"; tail -n 1 gpurun_out/bench_nograph.log | cut -c1-200
