#!/bin/bash
# Full GPU validation payload for one gpurun call: build check, every GPU test file in its own process (a trap in one
# kernel family cannot poison the others), smoke, bench (1 GPU) and the CPU reference arm.  Logs under gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c 'import __graft_entry__ as g; g.build()' > gpurun_out/build.log 2>&1; echo "build rc=$?"
for f in elementwise attention hog gemm modules graph ddp; do
  timeout 600 python -m pytest tests/test_gpu_$f.py -m gpu -q > gpurun_out/test_$f.log 2>&1; echo "test_$f rc=$?"
  tail -n 2 gpurun_out/test_$f.log | cut -c1-200
done
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -n 1 gpurun_out/bench_ref.log | cut -c1-300
