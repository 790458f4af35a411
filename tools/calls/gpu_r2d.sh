#!/bin/bash
# round-2 call D: rank-4 map layouts, residual tests, in-situ A/B of every switch, profiles, ncu, full bench
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
export VT_EXPERIMENTAL=1
SP_OK=0
for lay in 1 0; do
  for c in "spatial 1" "spatial 3"; do
    VT_RES_4D_PFIRST=$lay timeout 120 python tools/res_probe.py $c > gpurun_out/probe.log 2>&1; rc=$?
    line=$(grep -E 'rel err' gpurun_out/probe.log | tail -n 1)
    echo "probe [$c, p_first=$lay] rc=$rc : ${line:-$(grep -E 'Error|error' gpurun_out/probe.log | tail -n 1 | cut -c1-120)}"
  done
done
VT_RES_4D_PFIRST=1 timeout 120 python tools/res_probe.py spatial 1 > gpurun_out/probe.log 2>&1 && grep -q "rel err 1\.[0-9]*e-07" gpurun_out/probe.log && SP_OK=1
echo "SP_OK=$SP_OK"
if [ "$SP_OK" = "1" ]; then export VT_TMA_RES_SPATIAL=1; K_RES="residual"; else K_RES="residual and plain"; fi
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "$K_RES" > gpurun_out/test_gemm_res.log 2>&1; echo "test_gemm [$K_RES] rc=$?"; tail -n 4 gpurun_out/test_gemm_res.log | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "on_tma" > gpurun_out/test_gemm_gelu.log 2>&1; echo "test_gemm gelu/dgelu tma rc=$?"; tail -n 3 gpurun_out/test_gemm_gelu.log | cut -c1-250
unset VT_EXPERIMENTAL
ab() {  # label, env...
  label=$1; shift
  env "$@" timeout 600 python bench.py --no-others --no-baselines --steps 20 > gpurun_out/ab_$label.log 2>&1
  grep '^{' gpurun_out/ab_$label.log | tail -n 1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); r = d['roofline']
    print('AB $label: %.3f ms  %.1f clips/s  gemm %.3f ms  kernels %s' % (d['ms_per_step'], d['value'], r.get('gemm_ms_per_step') or -1, d.get('kernels_per_replay')))
except Exception as e:
    print('AB $label: no line', e)
"
}
ab base VT_NONE=1
ab colsum VT_COLSUM_WIDE=1
ab ln2 VT_LN_BWD_V2=1
ab tail VT_TAIL_UNITS=1
ab res VT_TMA_RES=1
ab res_sp VT_TMA_RES=1 VT_TMA_RES_SPATIAL=$SP_OK
ab all VT_COLSUM_WIDE=1 VT_LN_BWD_V2=1 VT_TMA_RES=1 VT_TMA_RES_SPATIAL=$SP_OK
ab all_tail VT_COLSUM_WIDE=1 VT_LN_BWD_V2=1 VT_TMA_RES=1 VT_TMA_RES_SPATIAL=$SP_OK VT_TAIL_UNITS=1
ab base2 VT_NONE=1
FEAT="VT_COLSUM_WIDE=1 VT_LN_BWD_V2=1 VT_TMA_RES=1 VT_TMA_RES_SPATIAL=$SP_OK"
timeout 300 python tools/profile_step.py torchprof > gpurun_out/torchprof_base.log 2>&1; echo "torchprof base rc=$?"
env $FEAT timeout 300 python tools/profile_step.py torchprof > gpurun_out/torchprof_feat.log 2>&1; echo "torchprof feat rc=$?"
VT_TMA_RES_SPATIAL=$SP_OK timeout 600 python tools/gemm_table.py quick > gpurun_out/gemm_table.log 2>&1; echo "gemm_table rc=$?"
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench (shipping defaults, all workloads) rc=$?"; tail -n 1 gpurun_out/bench.log | cut -c1-300
env $FEAT bash tools/ncu_step.sh
