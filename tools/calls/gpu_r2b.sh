#!/bin/bash
# round-2 call B: GEMM tests (new residual / tail paths), module tests, bench, per-shape GEMM table
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x > gpurun_out/test_gemm.log 2>&1; echo "test_gemm rc=$?"; tail -n 3 gpurun_out/test_gemm.log | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_graph.py tests/test_gpu_head.py tests/test_gpu_mvit_oracle.py -q -m gpu -s > gpurun_out/test_modules.log 2>&1; echo "test_modules rc=$?"; grep -E "passed|failed" gpurun_out/test_modules.log | tail -n 2 | cut -c1-300; grep -E "^FAILED|^ERROR" gpurun_out/test_modules.log | head -n 20 | cut -c1-250
timeout 600 python tools/gemm_table.py quick > gpurun_out/gemm_table.log 2>&1; echo "gemm_table rc=$?"
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench.log | cut -c1-400
timeout 300 python tools/profile_step.py torchprof > gpurun_out/torchprof.log 2>&1; echo "torchprof rc=$?"
