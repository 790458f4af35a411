#!/bin/bash
# One gpurun payload: build check, GPU unit tests (one process per file so a trap in one kernel family
# cannot poison the others), smoke, short bench.  Everything is logged under gpurun_out/.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c 'import __graft_entry__ as g; g.build()' > gpurun_out/build.log 2>&1; echo "build rc=$?"
for v in "128 128 64 0 0" "128 256 64 0 0 256" "256 256 256 0 0" "256 256 128 0 1" "256 256 128 1 0" "256 256 128 1 1" "200 136 72 0 0" "1024 512 2048 1 1 0 4"; do
  timeout 120 python tools/gemm_probe.py $v >> gpurun_out/gemm_probe.log 2>&1; echo "probe [$v] rc=$?" >> gpurun_out/gemm_probe.log
done
tail -n 40 gpurun_out/gemm_probe.log
for f in elementwise attention hog gemm modules; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -x -q -s > gpurun_out/test_$f.log 2>&1; echo "test_$f rc=$?"
  tail -n 15 gpurun_out/test_$f.log
done
timeout 600 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -n 3 gpurun_out/bench.log
