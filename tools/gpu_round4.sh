#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_graph.py -m gpu -q -x > gpurun_out/test_graph.log 2>&1; echo "test_graph rc=$?"
tail -n 15 gpurun_out/test_graph.log
timeout 900 python tools/gemm_table.py > gpurun_out/gemm_table.log 2>&1; echo "gemm_table rc=$?"; cat gpurun_out/gemm_table.log
