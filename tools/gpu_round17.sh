#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1d.csv python tools/profile_step.py ncu > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm2_tcgen05|gemm_tcgen05|attn_tc|attn8|ln_bwd" -c 40 -o gpurun_out/step_r1d python tools/profile_step.py ncu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -n 2 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
