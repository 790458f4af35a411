#!/bin/bash
# 2-GPU sanity: the data-parallel bench path (graph capture + bucketed NCCL all-reduce)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
P='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ("n_gpus","value","ms_per_step")}, d["e2e"]["value"])'
NCCL_DEBUG=WARN timeout 75 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?"; tail -n 1 gpurun_out/bench_n2.log | python -c "$P" || tail -n 20 gpurun_out/bench_n2.log | cut -c1-300
