#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
P='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ("n_gpus","value","ms_per_step")}, d["e2e"]["value"])'
timeout 150 python -m pytest tests/test_gpu_ddp.py -m gpu -q -x > gpurun_out/test_ddp.log 2>&1; echo "test_ddp rc=$?"; tail -n 2 gpurun_out/test_ddp.log | cut -c1-200
for N in 2 4; do
NCCL_DEBUG=WARN timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; echo "bench n$N rc=$?"; tail -n 1 gpurun_out/bench_n$N.log | python -c "$P"
done
