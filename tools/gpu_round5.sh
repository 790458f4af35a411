#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_elementwise.py -m gpu -q -x > gpurun_out/test_gemm.log 2>&1; echo "test_gemm+elementwise rc=$?"; tail -n 5 gpurun_out/test_gemm.log
timeout 600 python -m pytest tests/test_gpu_graph.py -m gpu -q -x > gpurun_out/test_graph.log 2>&1; echo "test_graph rc=$?"; tail -n 8 gpurun_out/test_graph.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_modules.py -m gpu -q > gpurun_out/test_modules.log 2>&1; echo "test_modules rc=$?"; tail -n 3 gpurun_out/test_modules.log
timeout 900 python tools/gemm_table.py > gpurun_out/gemm_table.log 2>&1; echo "gemm_table rc=$?"; cat gpurun_out/gemm_table.log | cut -c1-330
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-graph > gpurun_out/bench_nograph.log 2>&1; echo "bench nograph rc=$?"; tail -n 1 gpurun_out/bench_nograph.log | cut -c1-400
