#!/bin/bash
# round-2 call O: shipping build with narrow tail units on — whole GPU suite, smoke, bench line
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 3 gpurun_out/pytest_gpu_all.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/smoke.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench.log | cut -c1-400
timeout 600 python tools/maskfeat_bench.py --graph > gpurun_out/maskfeat_o.log 2>&1; grep "CUDA graph" gpurun_out/maskfeat_o.log | cut -c1-160
