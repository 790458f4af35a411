#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 120 python tools/attn_phases.py > gpurun_out/attn_phases.log 2>&1; echo "attn_phases rc=$?"; cat gpurun_out/attn_phases.log
timeout 200 python tools/profile_step.py torchprof > gpurun_out/torchprof.log 2>&1; echo "torchprof rc=$?"; sed -n 3,34p gpurun_out/torchprof.log | cut -c1-130
