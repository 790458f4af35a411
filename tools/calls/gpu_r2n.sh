#!/bin/bash
# round-2 call N: narrow tail units on the workloads whose every GEMM has a partial last row of tiles (ViViT: M = 12608)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
ab() {
  label=$1; wl=$2; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --no-others --no-baselines --steps 20 > gpurun_out/ab_$label.log 2>&1
  grep '^{' gpurun_out/ab_$label.log | tail -n 1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); r = d.get('roofline') or {}
    print('AB $label: %.3f ms  %.1f clips/s  gemm %s ms kernels %s' % (d['ms_per_step'], d['value'], r.get('gemm_ms_per_step'), d.get('kernels_per_replay')))
except Exception as e:
    print('AB $label: no line', e)
"
}
ab vivit_base vivit VT_NONE=1
ab vivit_tail vivit VT_TAIL_UNITS=1
ab vivit_base2 vivit VT_NONE=1
ab vivit_tail2 vivit VT_TAIL_UNITS=1
ab mvit_base mvit VT_NONE=1
ab mvit_tail mvit VT_TAIL_UNITS=1
ab ts_tail timesformer VT_TAIL_UNITS=1
ab ts_base timesformer VT_NONE=1
