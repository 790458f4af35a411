#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 300 python tools/graph_debug.py > gpurun_out/graph_debug.log 2>&1; echo "graph_debug rc=$?"; grep -v Warn gpurun_out/graph_debug.log | cut -c1-250 | head -30
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench n1 rc=$?"; tail -n 1 gpurun_out/bench_n1.log | cut -c1-230
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?"; tail -n 4 gpurun_out/bench_n2.log | cut -c1-400
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-graph > gpurun_out/bench_n2_nograph.log 2>&1; echo "bench n2 nograph rc=$?"; tail -n 2 gpurun_out/bench_n2_nograph.log | cut -c1-300
