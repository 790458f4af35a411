#!/bin/bash
# 2-GPU check of the shipping build: DDP test, bench line with ddp_check (buckets zeroed once per step inside the graph)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 600 python -m pytest tests/test_gpu_ddp.py -q -m gpu > gpurun_out/test_ddp.log 2>&1; echo "test_ddp rc=$?"; tail -n 2 gpurun_out/test_ddp.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_final.log 2>&1; echo "bench n2 rc=$?"
grep '^{' gpurun_out/bench_n2_final.log | tail -n 1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d.get(k) for k in ('value', 'ms_per_step')}, 'ddp_check', (d.get('ddp_check') or {}).get('rel_l2'), 'exchange', (d.get('exchange') or {}).get('allreduce_exposed_ms'))
print({k: (v.get('value'), v.get('ms_per_step'), v.get('error')) for k, v in (d.get('other_workloads') or {}).items()})
"
