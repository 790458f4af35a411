#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_attention.py -m gpu -q -x > gpurun_out/test_attention.log 2>&1; echo "test_attention rc=$?"
tail -n 25 gpurun_out/test_attention.log
timeout 900 python -m pytest tests/test_gpu_modules.py -m gpu -q -s > gpurun_out/test_modules.log 2>&1; echo "test_modules rc=$?"
grep -E "rel-L2|golden|layer:|vivit|passed|failed|Error" gpurun_out/test_modules.log | head -30
timeout 600 python tools/profile_step.py torchprof > gpurun_out/torchprof.log 2>&1; echo "torchprof rc=$?"; tail -n 32 gpurun_out/torchprof.log | cut -c1-150
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench.log
