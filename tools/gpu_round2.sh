#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_gpu_modules.py -m gpu -q -s > gpurun_out/test_modules.log 2>&1; echo "test_modules rc=$?"
grep -E "rel-L2|golden|layer:|grad |vivit|passed|failed|Error" gpurun_out/test_modules.log | head -60
timeout 600 python tools/profile_step.py torchprof > gpurun_out/torchprof.log 2>&1; echo "torchprof rc=$?"; cat gpurun_out/torchprof.log | tail -50
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1a.csv python tools/profile_step.py ncu > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
