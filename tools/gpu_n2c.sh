#!/bin/bash
# 2-GPU re-check after aligning the bucket slices
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 400 python -m pytest tests/test_gpu_ddp.py -q -m gpu > gpurun_out/test_ddp.log 2>&1; echo "test_ddp rc=$?"; tail -n 2 gpurun_out/test_ddp.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --no-others --no-baselines --steps 10 > gpurun_out/bench_n2_c.log 2>&1; echo "bench n2 rc=$?"
grep '^{' gpurun_out/bench_n2_c.log | tail -n 1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d.get(k) for k in ('value', 'ms_per_step')}, 'ddp_check', (d.get('ddp_check') or {}).get('rel_l2'), 'exchange', (d.get('exchange') or {}).get('allreduce_exposed_ms'))
"
