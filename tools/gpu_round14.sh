#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 90 python tools/gemm_probe.py 1000 768 768 0 0 > gpurun_out/probe.log 2>&1; echo "probe rc=$?"; grep -E "rel=|bad|Error" gpurun_out/probe.log | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/test_gemm.log 2>&1; echo "test_gemm rc=$?"; tail -n 4 gpurun_out/test_gemm.log | cut -c1-300
timeout 120 python tools/gemm_phases.py > gpurun_out/gemm_phases.log 2>&1; echo "phases rc=$?"; cat gpurun_out/gemm_phases.log
timeout 200 python tools/gemm_table.py > gpurun_out/gemm_table.log 2>&1; echo "gemm_table rc=$?"; cut -c1-210 gpurun_out/gemm_table.log
timeout 200 python -m pytest tests/test_gpu_modules.py -m gpu -q > gpurun_out/test_modules.log 2>&1; echo "test_modules rc=$?"; tail -n 2 gpurun_out/test_modules.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench.log | cut -c1-250
