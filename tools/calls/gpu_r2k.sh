#!/bin/bash
# round-2 call K: merged proj + temporal_fc (product weight), second bias, dual column sums
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_elementwise.py -q -m gpu -k "second_bias or column_sums" > gpurun_out/test_k1.log 2>&1; echo "bias2 + column sums rc=$?"; tail -n 3 gpurun_out/test_k1.log | cut -c1-200
VT_MERGE_TEMPORAL_FC=1 timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_baseline_shapes.py tests/test_gpu_graph.py tests/test_gpu_ddp.py -q -m gpu > gpurun_out/test_k2.log 2>&1; echo "module / baseline-shape / graph tests, merged fc rc=$?"; tail -n 5 gpurun_out/test_k2.log | cut -c1-250
grep -E "timesformer.*(feature|loss|grad)" gpurun_out/test_k2.log | head -n 5
ab() {
  label=$1; shift
  env "$@" timeout 600 python bench.py --no-others --no-baselines --steps 20 > gpurun_out/ab_$label.log 2>&1
  grep '^{' gpurun_out/ab_$label.log | tail -n 1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); r = d['roofline']
    print('AB $label: %.3f ms  %.1f clips/s  gemm %.3f ms frac %.3f kernels %s loss %s' % (d['ms_per_step'], d['value'], r.get('gemm_ms_per_step') or -1, r.get('frac') or -1, d.get('kernels_per_replay'), d.get('loss')))
except Exception as e:
    print('AB $label: no line', e)
"
}
ab dflt VT_NONE=1
ab merged VT_MERGE_TEMPORAL_FC=1
ab dflt2 VT_NONE=1
ab merged2 VT_MERGE_TEMPORAL_FC=1
VT_MERGE_TEMPORAL_FC=1 timeout 300 python tools/profile_step.py torchprof > gpurun_out/torchprof_merged.log 2>&1; echo "torchprof rc=$?"
head -n 30 gpurun_out/torchprof_merged.log | tail -n 24 | cut -c1-150
du -sh gpurun_out
