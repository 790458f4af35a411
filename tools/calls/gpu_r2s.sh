#!/bin/bash
# round-2 call S: BASELINE-shape parity numbers of the shipping build (printed), for the docs
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_gpu_baseline_shapes.py -q -m gpu -s 2>&1 | grep -E "rel-L2|parameter gradients|passed|failed" | cut -c1-220
