#!/bin/bash
# round-2 call I: residual epilogue with row-by-row straddling groups (all maps), producer column sums with a parallel final pass
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
rm -f gpurun_out/*.ncu-rep
export VT_EXPERIMENTAL=1
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "residual and plain" > gpurun_out/test_res_plain.log 2>&1; echo "residual plain rc=$?"; tail -n 2 gpurun_out/test_res_plain.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "residual and maps" > gpurun_out/test_res_maps.log 2>&1; rcm=$?; echo "residual maps rc=$rcm"; tail -n 6 gpurun_out/test_res_maps.log | cut -c1-200
unset VT_EXPERIMENTAL
SP=0; [ "$rcm" = "0" ] && SP=1
for c in 1 3; do
  VT_TMA_RES_SPATIAL=1 timeout 120 python tools/res_probe.py spatial $c > gpurun_out/probe_sp$c.log 2>&1
  echo "probe spatial cluster=$c: $(grep -E 'rel err|CUDA error' gpurun_out/probe_sp$c.log | tail -n 1 | cut -c1-160)"
done
timeout 900 python -m pytest tests/test_gpu_elementwise.py -q -m gpu -k "column_sums" > gpurun_out/test_fc.log 2>&1; echo "fused colsum tests rc=$?"; tail -n 2 gpurun_out/test_fc.log | cut -c1-200
VT_FUSED_COLSUM=1 VT_TMA_RES_SPATIAL=$SP timeout 900 python -m pytest tests/test_gpu_modules.py tests/test_gpu_baseline_shapes.py tests/test_gpu_graph.py -q -m gpu -x > gpurun_out/test_modules_fused.log 2>&1; echo "module + baseline-shape + graph tests (fused colsum, spatial=$SP) rc=$?"; tail -n 3 gpurun_out/test_modules_fused.log | cut -c1-250
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
ab fused_sp VT_FUSED_COLSUM=1 VT_TMA_RES_SPATIAL=$SP
ab old2 VT_TMA_RES=0 VT_ROWS_SPLIT=0
ab fused_sp2 VT_FUSED_COLSUM=1 VT_TMA_RES_SPATIAL=$SP
VT_FUSED_COLSUM=1 VT_TMA_RES_SPATIAL=$SP timeout 300 python tools/profile_step.py torchprof > gpurun_out/torchprof_fused.log 2>&1; echo "torchprof rc=$?"
grep -E "colsum|gather_cast|gelu_bwd|reduce_rows" gpurun_out/torchprof_fused.log | cut -c1-150
for f in 0 1; do
  VT_FUSED_COLSUM=$f timeout 600 python tools/maskfeat_bench.py --graph > gpurun_out/maskfeat_f$f.log 2>&1; echo "maskfeat fused_colsum=$f: $(grep 'CUDA graph' gpurun_out/maskfeat_f$f.log | cut -c1-150)"
done
du -sh gpurun_out
