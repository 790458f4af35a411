#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python -m pytest tests/test_gpu_gemm.py -m gpu -q -x > gpurun_out/test_gemm.log 2>&1; echo "test_gemm rc=$?"; tail -n 2 gpurun_out/test_gemm.log | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_ddp.py -m gpu -q -x > gpurun_out/test_ddp.log 2>&1; echo "test_ddp rc=$?"; tail -n 6 gpurun_out/test_ddp.log | cut -c1-300
timeout 200 python tools/vivit_bench.py > gpurun_out/vivit.log 2>&1; echo "vivit rc=$?"; tail -n 3 gpurun_out/vivit.log
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench n1 rc=$?"; tail -n 1 gpurun_out/bench_n1.log | cut -c1-260
NCCL_DEBUG=WARN timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?"; tail -n 1 gpurun_out/bench_n2.log | cut -c1-260
