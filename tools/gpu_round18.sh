#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py -m gpu -q -x > gpurun_out/test_k.log 2>&1; echo "test gemm+attn rc=$?"; tail -n 3 gpurun_out/test_k.log | cut -c1-200
timeout 200 python -m pytest tests/test_gpu_modules.py tests/test_gpu_graph.py -m gpu -q > gpurun_out/test_modules.log 2>&1; echo "test_modules rc=$?"; tail -n 2 gpurun_out/test_modules.log
timeout 200 python tools/profile_step.py torchprof > gpurun_out/torchprof.log 2>&1; echo "torchprof rc=$?"; sed -n 3,30p gpurun_out/torchprof.log | cut -c1-120
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench.log
