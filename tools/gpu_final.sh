#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu (single process) rc=$?"; tail -n 3 gpurun_out/pytest_gpu_all.log | cut -c1-200
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -n 1 gpurun_out/bench_ref.log | cut -c1-400
