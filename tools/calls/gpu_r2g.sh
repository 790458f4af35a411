#!/bin/bash
# round-2 call G: remainder-row split timing + tests, plane-major pooling adjoint, reworked wide column sums
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
rm -f gpurun_out/*.ncu-rep
timeout 300 python tools/rows_probe.py > gpurun_out/rows_probe.log 2>&1; echo "rows_probe rc=$?"; grep -E "us \|" gpurun_out/rows_probe.log | cut -c1-220
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "remainder" > gpurun_out/test_rows.log 2>&1; echo "test remainder rc=$?"; tail -n 4 gpurun_out/test_rows.log | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_mvit.py tests/test_gpu_mvit_oracle.py tests/test_gpu_elementwise.py -q -m gpu -k "pool or colsum" > gpurun_out/test_pool.log 2>&1; echo "pool+colsum tests rc=$?"; tail -n 4 gpurun_out/test_pool.log | cut -c1-250
for cfg in "0 0" "1 0" "1 1"; do
  set -- $cfg
  VT_POOL_V2=$1 VT_COLSUM_WIDE=$2 VT_ROWS_SPLIT=$2 VT_TMA_RES=$2 timeout 600 python tools/maskfeat_bench.py --graph --profile > gpurun_out/maskfeat_p$1x$2.log 2>&1
  echo "maskfeat pool_v2=$1 colsum+rows+res=$2 rc=$?: $(grep 'CUDA graph' gpurun_out/maskfeat_p$1x$2.log | cut -c1-150)"
done
grep -E "kernel time total|pool_|colsum|reduce_rows|rows_n" gpurun_out/maskfeat_p1x1.log | head -n 16 | cut -c1-150
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
ab colsum VT_COLSUM_WIDE=1
ab rows VT_ROWS_SPLIT=1
ab all VT_COLSUM_WIDE=1 VT_ROWS_SPLIT=1 VT_TMA_RES=1
ab base2 VT_NONE=1
ab all2 VT_COLSUM_WIDE=1 VT_ROWS_SPLIT=1 VT_TMA_RES=1
VT_COLSUM_WIDE=1 VT_ROWS_SPLIT=1 VT_TMA_RES=1 timeout 300 python tools/profile_step.py torchprof > gpurun_out/torchprof_all.log 2>&1; echo "torchprof rc=$?"
grep -E "colsum|rows_n|gemm" gpurun_out/torchprof_all.log | head -n 12 | cut -c1-150
du -sh gpurun_out
