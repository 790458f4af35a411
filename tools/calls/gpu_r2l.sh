#!/bin/bash
# round-2 call L: wider tall reduction kernel
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_mvit.py -q -m gpu > gpurun_out/test_l.log 2>&1; echo "elementwise + mvit tests rc=$?"; tail -n 3 gpurun_out/test_l.log | cut -c1-200
for i in 1 2; do
timeout 600 python bench.py --no-others --no-baselines --steps 20 > gpurun_out/ab_l$i.log 2>&1
grep '^{' gpurun_out/ab_l$i.log | tail -n 1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('AB shipping: %.3f ms  %.1f clips/s  gemm %.3f ms frac %.3f kernels %s' % (d['ms_per_step'], d['value'], r.get('gemm_ms_per_step') or -1, r.get('frac') or -1, d.get('kernels_per_replay')))
"
done
timeout 300 python tools/profile_step.py torchprof > gpurun_out/torchprof_l.log 2>&1; grep -E "reduce_rows|colsum|gather_cast" gpurun_out/torchprof_l.log | cut -c1-140
