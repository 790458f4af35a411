#!/bin/bash
# round-2 call U: the two switches that are still off, against the final build
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
ab() {
  label=$1; shift
  env "$@" timeout 600 python bench.py --no-others --no-baselines --steps 20 > gpurun_out/ab_u_$label.log 2>&1
  grep '^{' gpurun_out/ab_u_$label.log | tail -n 1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d.get('roofline') or {}
print('AB $label: %.3f ms  %.1f clips/s  gemm %s ms kernels %s' % (d['ms_per_step'], d['value'], r.get('gemm_ms_per_step'), d.get('kernels_per_replay')))
"
}
ab base VT_NONE=1
ab colsum_wide VT_COLSUM_WIDE=1
ab ln_v2 VT_LN_BWD_V2=1
ab both VT_COLSUM_WIDE=1 VT_LN_BWD_V2=1
ab base2 VT_NONE=1
