#!/bin/bash
# round-2 call J: release candidate — whole GPU suite, smoke, full bench line, MaskFeat profile, ncu evidence
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest -m gpu rc=$?"; tail -n 6 gpurun_out/pytest_gpu_all.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench.log | cut -c1-3000
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench reference rc=$?"; tail -n 1 gpurun_out/bench_ref.log | cut -c1-600
timeout 600 python tools/maskfeat_bench.py --graph --profile > gpurun_out/maskfeat_bench.log 2>&1; echo "maskfeat rc=$?"; grep -E "CUDA graph|kernel time total" gpurun_out/maskfeat_bench.log | cut -c1-160; grep -E "^ +[0-9.]+ ms" gpurun_out/maskfeat_bench.log | head -n 22 | cut -c1-130
timeout 300 python tools/profile_step.py torchprof > gpurun_out/torchprof_final.log 2>&1; echo "torchprof rc=$?"
bash tools/ncu_step.sh
du -sh gpurun_out
