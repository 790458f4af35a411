#!/bin/bash
# round-2 call H: bias gradients from the dY producers; defaults now: TMA residual (plain + temporal), remainder rows, pooling gen 2
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
rm -f gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_gemm.py -q -m gpu -k "column_sums or remainder or colsum" > gpurun_out/test_h.log 2>&1; echo "fused colsum + remainder tests rc=$?"; tail -n 4 gpurun_out/test_h.log | cut -c1-250
ab() {
  label=$1; shift
  env "$@" timeout 600 python bench.py --no-others --no-baselines --steps 20 > gpurun_out/ab_$label.log 2>&1
  grep '^{' gpurun_out/ab_$label.log | tail -n 1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); r = d['roofline']
    print('AB $label: %.3f ms  %.1f clips/s  gemm %.3f ms frac %.3f kernels %s' % (d['ms_per_step'], d['value'], r.get('gemm_ms_per_step') or -1, r.get('frac') or -1, d.get('kernels_per_replay')))
except Exception as e:
    print('AB $label: no line', e)
"
}
ab old VT_TMA_RES=0 VT_ROWS_SPLIT=0
ab dflt VT_NONE=1
ab fused VT_FUSED_COLSUM=1
ab old2 VT_TMA_RES=0 VT_ROWS_SPLIT=0
ab fused2 VT_FUSED_COLSUM=1
VT_FUSED_COLSUM=1 timeout 300 python tools/profile_step.py torchprof > gpurun_out/torchprof_fused.log 2>&1; echo "torchprof rc=$?"
head -n 34 gpurun_out/torchprof_fused.log | tail -n 29 | cut -c1-150
VT_FUSED_COLSUM=1 timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_baseline_shapes.py -q -m gpu -x > gpurun_out/test_modules_fused.log 2>&1; echo "module + baseline-shape tests with fused colsum rc=$?"; tail -n 3 gpurun_out/test_modules_fused.log | cut -c1-250
for f in 0 1; do
  VT_FUSED_COLSUM=$f timeout 600 python tools/maskfeat_bench.py --graph > gpurun_out/maskfeat_f$f.log 2>&1; echo "maskfeat fused_colsum=$f: $(grep 'CUDA graph' gpurun_out/maskfeat_f$f.log | cut -c1-150)"
done
du -sh gpurun_out
