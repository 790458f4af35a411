#!/bin/bash
# round-2 call Q: cls-row kernel
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_modules.py tests/test_gpu_graph.py tests/test_gpu_baseline_shapes.py -q -m gpu -k "cls_rows or timesformer or vivit or graph or baseline" > gpurun_out/test_q.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/test_q.log | cut -c1-200
for i in 1 2; do
timeout 600 python bench.py --no-others --no-baselines --steps 20 > gpurun_out/ab_q$i.log 2>&1
grep '^{' gpurun_out/ab_q$i.log | tail -n 1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d.get('roofline') or {}
print('timesformer: %.3f ms  %.1f clips/s  gemm %s ms kernels %s loss %s' % (d['ms_per_step'], d['value'], r.get('gemm_ms_per_step'), d.get('kernels_per_replay'), d.get('loss')))
"
done
