#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
P='import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("timing"))'
timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench n1 rc=$?"; tail -n 1 gpurun_out/bench_n1.log | python -c "$P"; grep -i "roofline:" gpurun_out/bench_n1.log | head -3
for R in 0 8 16; do
NCCL_DEBUG=WARN timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((R%10)) bench.py --gpus 2 --steps 20 --warmup 3 --reserve-sms $R > gpurun_out/bench_n2_r$R.log 2>&1; echo "bench n2 reserve=$R rc=$?"; tail -n 1 gpurun_out/bench_n2_r$R.log | python -c "$P"
done
