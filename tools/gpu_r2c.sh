#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
export CUDA_LAUNCH_BLOCKING=1
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_gpu_gemm.py::test_residual_epilogue_tma_plain_rows[1000-768-256-128-single-cta]" -q -m gpu -x > gpurun_out/sanitizer_single.log 2>&1; echo "sanitizer single rc=$?"
grep -E "=========|Error|error" gpurun_out/sanitizer_single.log | head -n 40 | cut -c1-300
timeout 600 python -m pytest "tests/test_gpu_gemm.py::test_residual_epilogue_tma_plain_rows[1000-768-256-256-single-cta]" -q -m gpu -x > gpurun_out/t256.log 2>&1; echo "bn256 single rc=$?"; tail -n 3 gpurun_out/t256.log | cut -c1-200
timeout 600 python -m pytest "tests/test_gpu_gemm.py::test_residual_epilogue_tma_plain_rows[4096-256-64-128-single-cta]" -q -m gpu -x > gpurun_out/t4096.log 2>&1; echo "M4096 (no partial tiles) single rc=$?"; tail -n 3 gpurun_out/t4096.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "narrow_tail and not residual" > gpurun_out/tail.log 2>&1; echo "tail rc=$?"; tail -n 5 gpurun_out/tail.log | cut -c1-200
