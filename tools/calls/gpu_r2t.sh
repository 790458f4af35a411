#!/bin/bash
# round-2 call T: last check of the committed tree — GPU suite, smoke, driver-style bench
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 2 gpurun_out/pytest_gpu_all.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log | cut -c1-200
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench.log | cut -c1-330
