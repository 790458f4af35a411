#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
P='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ("n_gpus","value","ms_per_step")}, d["e2e"]["value"])'
NCCL_DEBUG=WARN timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.log 2>&1; echo "bench n8 rc=$?"; tail -n 1 gpurun_out/bench_n8.log | python -c "$P" || tail -n 20 gpurun_out/bench_n8.log | cut -c1-300
timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_ref_n8.log 2>&1; echo "ref n8 rc=$?"; tail -n 1 gpurun_out/bench_ref_n8.log | cut -c1-200
