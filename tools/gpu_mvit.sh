#!/bin/bash
# One-shot GPU validation of the MaskFeat / MViT path: kernel + module tests, then the full-size step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader > gpurun_out/gpu.txt
timeout 420 python -m pytest tests/test_gpu_mvit.py -q -m gpu -p no:cacheprovider 2>&1 | tail -80 > gpurun_out/test_mvit.log
echo "pytest rc=${PIPESTATUS[0]}"; tail -25 gpurun_out/test_mvit.log
timeout 240 python tools/maskfeat_bench.py --batch 8 --steps 3 --warmup 2 --profile --graph > gpurun_out/maskfeat_bench.log 2>&1
echo "bench rc=$?"; tail -40 gpurun_out/maskfeat_bench.log
