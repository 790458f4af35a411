#!/bin/bash
# 2-GPU call: DDP correctness (pytest + ddp_check in the bench line), exchange cost, direct bucket writes on / off
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
FEAT="${VT_FEATURES:-}"
timeout 600 python -m pytest tests/test_gpu_ddp.py -q -m gpu > gpurun_out/test_ddp.log 2>&1; echo "test_ddp rc=$?"; tail -n 2 gpurun_out/test_ddp.log | cut -c1-200
run() {   # name, extra env ..., then bench args after --
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env $FEAT "${envs[@]}" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 "$@" > gpurun_out/bench_n2_$name.log 2>&1; echo "bench n2 $name rc=$?"
  grep '^{' gpurun_out/bench_n2_$name.log | tail -n 1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print({k: d.get(k) for k in ('value', 'ms_per_step', 'ddp_check', 'exchange')}, 'gemm_ms', d['roofline'].get('gemm_ms_per_step'))
    ow = d.get('other_workloads') or {}
    print({k: (v.get('value'), v.get('ms_per_step'), v.get('error')) for k, v in ow.items()})
except Exception as e:
    print('no line', e)
"
}
run copy VT_DDP_DIRECT=0 -- --no-others --steps 10
run direct VT_DDP_DIRECT=1 -- --no-others --steps 10
run direct_r8 VT_DDP_DIRECT=1 -- --no-others --steps 10 --reserve-sms 8
run full VT_DDP_DIRECT=0 -- --steps 5
