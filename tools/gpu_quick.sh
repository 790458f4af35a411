#!/bin/bash
# quick check: GEMM/elementwise kernel tests + the bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_elementwise.py tests/test_gpu_modules.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout 300 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench.log | cut -c1-330
