#!/bin/bash
# round-2 call F: spatial residual boxes with element stride, remainder-row split, MaskFeat profile with the new pooling kernels
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
rm -f gpurun_out/*.ncu-rep
SP_OK=0
for c in 1 3; do
  timeout 120 python tools/res_probe.py spatial $c > gpurun_out/probe_sp$c.log 2>&1; rc=$?
  echo "probe spatial cluster=$c rc=$rc: $(grep -E 'rel err|CUDA error' gpurun_out/probe_sp$c.log | tail -n 1 | cut -c1-160)"
done
grep -q "rel err [0-9.]*e-0[6-9]" gpurun_out/probe_sp1.log && grep -q "rel err [0-9.]*e-0[6-9]" gpurun_out/probe_sp3.log && SP_OK=1
echo "SP_OK=$SP_OK"
export VT_EXPERIMENTAL=1
[ "$SP_OK" = "1" ] && export VT_TMA_RES_SPATIAL=1
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "residual or remainder" > gpurun_out/test_gemm_f.log 2>&1; echo "test_gemm residual+remainder rc=$?"; tail -n 5 gpurun_out/test_gemm_f.log | cut -c1-250
unset VT_EXPERIMENTAL
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
ab base VT_NONE=1
ab res VT_TMA_RES=1
ab res_sp VT_TMA_RES=1 VT_TMA_RES_SPATIAL=$SP_OK
ab res_sp_rows VT_TMA_RES=1 VT_TMA_RES_SPATIAL=$SP_OK VT_ROWS_SPLIT=1
ab res_sp_rows_cs VT_TMA_RES=1 VT_TMA_RES_SPATIAL=$SP_OK VT_ROWS_SPLIT=1 VT_COLSUM_WIDE=1
ab rows VT_ROWS_SPLIT=1
ab base2 VT_NONE=1
VT_TMA_RES=1 VT_TMA_RES_SPATIAL=$SP_OK VT_ROWS_SPLIT=1 timeout 600 python tools/gemm_table.py quick > gpurun_out/gemm_table_f.log 2>&1; echo "gemm_table rc=$?"
for cfg in "0 0 0" "1 0 0" "1 1 1"; do
  set -- $cfg
  VT_POOL_V2=$1 VT_TMA_RES=$2 VT_ROWS_SPLIT=$3 VT_COLSUM_WIDE=$3 timeout 600 python tools/maskfeat_bench.py --graph --profile > gpurun_out/maskfeat_p$1r$2s$3.log 2>&1
  echo "maskfeat pool_v2=$1 res=$2 rows+colsum=$3 rc=$?: $(grep 'CUDA graph' gpurun_out/maskfeat_p$1r$2s$3.log | cut -c1-150)"
done
grep -E "kernel time total|pool_|ln_small|xattn|colsum|reduce_rows|gemm" gpurun_out/maskfeat_p1r1s1.log | head -n 30 | cut -c1-150
du -sh gpurun_out
