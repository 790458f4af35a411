#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 120 python tools/gemm_phases.py > gpurun_out/gemm_phases.log 2>&1; echo "phases rc=$?"; cat gpurun_out/gemm_phases.log
