#!/bin/bash
# round-2 call P: gradient arena (one memset per step instead of one per split-K weight-gradient GEMM)
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_graph.py tests/test_gpu_baseline_shapes.py -q -m gpu -k "split_k or graph or baseline or timesformer" > gpurun_out/test_p.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/test_p.log | cut -c1-200
for i in 1 2; do
for wl in timesformer; do
timeout 600 python bench.py --workload $wl --no-others --no-baselines --steps 20 > gpurun_out/ab_p_$wl$i.log 2>&1
grep '^{' gpurun_out/ab_p_$wl$i.log | tail -n 1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d.get('roofline') or {}
print('$wl: %.3f ms  %.1f clips/s  gemm %s ms kernels %s loss %s' % (d['ms_per_step'], d['value'], r.get('gemm_ms_per_step'), d.get('kernels_per_replay'), d.get('loss')))
"
done
done
for wl in vivit mvit maskfeat; do
timeout 600 python bench.py --workload $wl --no-others --no-baselines --steps 10 > gpurun_out/ab_p_$wl.log 2>&1
grep '^{' gpurun_out/ab_p_$wl.log | tail -n 1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d.get('roofline') or {}
print('$wl: %.3f ms  %.1f clips/s  kernels %s loss %s' % (d['ms_per_step'], d['value'], d.get('kernels_per_replay'), d.get('loss')))
"
done
