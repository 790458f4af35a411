#!/bin/bash
# 8-GPU call: scaling of the shipping step, exchange cost, ddp_check with the merged temporal GEMM; NCCL channel count A/B
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
run() {
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 8 "$@" > gpurun_out/bench_n8_$name.log 2>&1; echo "bench n8 $name rc=$?"
  grep '^{' gpurun_out/bench_n8_$name.log | tail -n 1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print({k: d.get(k) for k in ('value', 'ms_per_step')}, 'ddp_check', (d.get('ddp_check') or {}).get('rel_l2'), 'exchange', (d.get('exchange') or {}).get('allreduce_exposed_ms'), 'gemm_ms', d['roofline'].get('gemm_ms_per_step'))
except Exception as e:
    print('no line', e)
"
}
run dflt VT_NONE=1 -- --no-others --no-baselines --steps 10
run nch8 NCCL_MAX_NCHANNELS=8 -- --no-others --no-baselines --steps 10
tail -n 3 gpurun_out/bench_n8_dflt.log | cut -c1-300
