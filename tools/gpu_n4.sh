#!/bin/bash
# 4-GPU line of the shipping build
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --no-others --no-baselines --steps 10 > gpurun_out/bench_n4.log 2>&1; echo "bench n4 rc=$?"
grep '^{' gpurun_out/bench_n4.log | tail -n 1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d.get(k) for k in ('value', 'ms_per_step')}, 'ddp_check', (d.get('ddp_check') or {}).get('rel_l2'), 'exchange', (d.get('exchange') or {}).get('allreduce_exposed_ms'))
"
