#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 200 python -m pytest tests/test_gpu_graph.py -m gpu -q > gpurun_out/test_graph.log 2>&1; echo "test_graph rc=$?"; tail -n 2 gpurun_out/test_graph.log | cut -c1-200
NCCL_DEBUG=WARN timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?"; tail -n 1 gpurun_out/bench_n2.log | cut -c1-1200
NCCL_DEBUG=WARN timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-graph > gpurun_out/bench_n2_nograph.log 2>&1; echo "bench n2 nograph rc=$?"; tail -n 1 gpurun_out/bench_n2_nograph.log | cut -c1-300
