#!/bin/bash
# round-2 call C: validate the switched paths one by one, then measure (bench, GEMM table, profile, ncu) with every path that passed
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
export VT_EXPERIMENTAL=1
: > gpurun_out/features_ok.txt
# ---- 1. TMA residual epilogue: one case per process
RES_OK=1
for c in "plain 1" "plain 3" "plain 1 1000" "plain 3 12552" "temporal 1" "temporal 3" "spatial 1" "spatial 3"; do
  timeout 120 python tools/res_probe.py $c > gpurun_out/probe.log 2>&1; rc=$?
  line=$(grep -E 'rel err' gpurun_out/probe.log | tail -n 1)
  echo "probe [$c] rc=$rc : ${line:-$(grep -E 'Error|error' gpurun_out/probe.log | tail -n 1 | cut -c1-160)}"
  ok=$(python -c "import sys; l='''$line'''; print(1 if l and float(l.split()[-1]) < 1e-4 else 0)" 2>/dev/null || echo 0)
  if [ "$ok" != "1" ]; then RES_OK=0; fi
done
echo "RES_OK=$RES_OK"
if [ "$RES_OK" != "1" ]; then
  T1="tests/test_gpu_gemm.py::test_residual_epilogue_tma_plain_rows[4096-256-64-128-single-cta]"
  timeout 600 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest "$T1" -q -m gpu -x > gpurun_out/sanitizer_t1.log 2>&1; echo "sanitizer T1 rc=$?"
  grep -E "=========" gpurun_out/sanitizer_t1.log | head -n 30 | cut -c1-260
else
  timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "residual" > gpurun_out/test_gemm_res.log 2>&1; rc=$?; echo "test_gemm residual rc=$rc"; tail -n 3 gpurun_out/test_gemm_res.log | cut -c1-250
  if [ "$rc" != "0" ]; then RES_OK=0; fi
fi
# ---- 2. everything else (experimental tests on)
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "not residual and not narrow_tail and not on_tma" > gpurun_out/test_gemm_base.log 2>&1; echo "test_gemm base rc=$?"; tail -n 2 gpurun_out/test_gemm_base.log | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "narrow_tail" > gpurun_out/test_gemm_tail.log 2>&1; TAIL_RC=$?; echo "test_gemm tail rc=$TAIL_RC"; tail -n 4 gpurun_out/test_gemm_tail.log | cut -c1-250
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "on_tma" > gpurun_out/test_gemm_gelu.log 2>&1; GELU_RC=$?; echo "test_gemm gelu/dgelu tma rc=$GELU_RC"; tail -n 4 gpurun_out/test_gemm_gelu.log | cut -c1-250
timeout 600 python -m pytest tests/test_gpu_elementwise.py -q -m gpu > gpurun_out/test_elem.log 2>&1; COLSUM_RC=$?; echo "test_elementwise (wide colsum on) rc=$COLSUM_RC"; tail -n 3 gpurun_out/test_elem.log | cut -c1-250
VT_LN_BWD_V2=1 timeout 600 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_mvit.py -q -m gpu -k "layernorm or ln" > gpurun_out/test_ln2.log 2>&1; LN_RC=$?; echo "LN bwd v2 rc=$LN_RC"; tail -n 2 gpurun_out/test_ln2.log | cut -c1-200
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_gemm.py --deselect tests/test_gpu_elementwise.py > gpurun_out/pytest_gpu_rest.log 2>&1; echo "pytest -m gpu rest rc=$?"
grep -E "passed|failed" gpurun_out/pytest_gpu_rest.log | tail -n 2 | cut -c1-300; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_rest.log | head -n 20 | cut -c1-250
FEAT=""
[ "$RES_OK" = "1" ] && FEAT="$FEAT VT_TMA_RES=1"
[ "$TAIL_RC" = "0" ] && FEAT="$FEAT VT_TAIL_UNITS=1"
[ "$COLSUM_RC" = "0" ] && FEAT="$FEAT VT_COLSUM_WIDE=1"
[ "$LN_RC" = "0" ] && FEAT="$FEAT VT_LN_BWD_V2=1"
echo "features that passed:$FEAT" | tee gpurun_out/features_ok.txt
[ "$GELU_RC" = "0" ] && echo "gelu-tma passed" >> gpurun_out/features_ok.txt
# ---- 3. measurements
unset VT_EXPERIMENTAL
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench (shipping defaults) rc=$?"; tail -n 1 gpurun_out/bench.log | cut -c1-700
env $FEAT timeout 600 python bench.py --no-others > gpurun_out/bench_feat.log 2>&1; echo "bench ($FEAT) rc=$?"; tail -n 1 gpurun_out/bench_feat.log | cut -c1-300
if [ "$GELU_RC" = "0" ]; then
  env $FEAT VT_TMA_GELU=1 VT_FUSED_GELU=1 timeout 600 python bench.py --no-others > gpurun_out/bench_feat_gelu.log 2>&1; echo "bench (+fused gelu) rc=$?"; tail -n 1 gpurun_out/bench_feat_gelu.log | cut -c1-300
  env $FEAT VT_TMA_GELU=1 VT_FUSED_GELU=1 VT_TMA_DGELU=1 timeout 600 python bench.py --no-others > gpurun_out/bench_feat_gelu2.log 2>&1; echo "bench (+fused gelu +dgelu) rc=$?"; tail -n 1 gpurun_out/bench_feat_gelu2.log | cut -c1-300
fi
VT_TMA_RES=$RES_OK timeout 600 python tools/gemm_table.py quick > gpurun_out/gemm_table.log 2>&1; echo "gemm_table rc=$?"
env $FEAT timeout 300 python tools/profile_step.py torchprof > gpurun_out/torchprof.log 2>&1; echo "torchprof rc=$?"
timeout 600 python tools/maskfeat_bench.py --graph --profile > gpurun_out/maskfeat_bench.log 2>&1; echo "maskfeat_bench rc=$?"
env $FEAT bash tools/ncu_step.sh
