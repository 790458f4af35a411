#!/bin/bash
# round-2 call A: whole GPU suite (no -x: see every failure), smoke, bench, kernel-time breakdown of the step
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu rc=$?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu_all.log | tail -n 3 | cut -c1-300
grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_all.log | head -n 30 | cut -c1-250
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log | cut -c1-200
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench.log | cut -c1-1500
timeout 300 python tools/profile_step.py torchprof > gpurun_out/torchprof.log 2>&1; echo "torchprof rc=$?"
