#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
rm -f gpurun_out/probe.log
for v in "256 256 128 0 0 256 0 3" "256 128 64 0 0 128 0 3" "1000 768 768 0 0 0 0 3" "512 512 256 1 1 256 0 3" "2304 768 12544 1 1 256 4 3" "12544 2304 768 0 0 256 0 3"; do
  timeout 90 python tools/gemm_probe.py $v >> gpurun_out/probe.log 2>&1; echo "probe [$v] rc=$?" >> gpurun_out/probe.log
done
grep -E "rel=|rc=|bad|matches|Error|error" gpurun_out/probe.log | cut -c1-200
timeout 240 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x -k "cluster or pair" > gpurun_out/test_gemm.log 2>&1; echo "test_gemm rc=$?"; tail -n 4 gpurun_out/test_gemm.log | cut -c1-300
timeout 200 python tools/gemm_table.py > gpurun_out/gemm_table.log 2>&1; echo "gemm_table rc=$?"; cut -c1-210 gpurun_out/gemm_table.log
