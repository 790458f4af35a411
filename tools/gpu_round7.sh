#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_graph.py -m gpu -q > gpurun_out/test_gemm.log 2>&1; echo "test_gemm+graph rc=$?"; tail -n 4 gpurun_out/test_gemm.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_modules.py -m gpu -q > gpurun_out/test_modules.log 2>&1; echo "test_modules rc=$?"; tail -n 3 gpurun_out/test_modules.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tcgen05 -o gpurun_out/gemm_r1b python tools/ncu_gemm_target.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_gemm.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench.log | cut -c1-230
