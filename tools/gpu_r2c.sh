#!/bin/bash
# round-2 call C: diagnose the TMA residual epilogue (sanitizer), validate everything else with it switched off
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
export VT_EXPERIMENTAL=1
for c in "plain 1" "plain 3" "plain 1 1000" "temporal 1" "temporal 3" "spatial 1" "spatial 3"; do
  timeout 120 python tools/res_probe.py $c > gpurun_out/probe.log 2>&1; echo "probe [$c] rc=$? : $(grep -E 'rel err|Error' gpurun_out/probe.log | tail -n 1 | cut -c1-160)"
done
T1="tests/test_gpu_gemm.py::test_residual_epilogue_tma_plain_rows[4096-256-64-128-single-cta]"
T2="tests/test_gpu_gemm.py::test_residual_epilogue_tma_plain_rows[1000-768-256-128-single-cta]"
timeout 300 python -m pytest "$T1" -q -m gpu -x > gpurun_out/res_t1.log 2>&1; echo "RES M4096 (whole tiles only) single rc=$?"; tail -n 2 gpurun_out/res_t1.log | cut -c1-200
timeout 300 python -m pytest "$T2" -q -m gpu -x > gpurun_out/res_t2.log 2>&1; echo "RES M1000 single rc=$?"; tail -n 2 gpurun_out/res_t2.log | cut -c1-200
timeout 600 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest "$T1" -q -m gpu -x > gpurun_out/sanitizer_t1.log 2>&1; echo "sanitizer T1 rc=$?"
grep -E "=========" gpurun_out/sanitizer_t1.log | head -n 40 | cut -c1-260
timeout 900 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "not residual" > gpurun_out/test_gemm_nores.log 2>&1; echo "test_gemm (all but residual, experimental on) rc=$?"; tail -n 12 gpurun_out/test_gemm_nores.log | cut -c1-250
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_gemm.py > gpurun_out/pytest_gpu_nores.log 2>&1; echo "pytest -m gpu (experimental on, w/o gemm file) rc=$?"
grep -E "passed|failed" gpurun_out/pytest_gpu_nores.log | tail -n 2 | cut -c1-300; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_nores.log | head -n 20 | cut -c1-250
VT_TMA_RES=0 timeout 600 python tools/gemm_table.py quick > gpurun_out/gemm_table_nores.log 2>&1; echo "gemm_table (generic residual epilogue) rc=$?"
timeout 600 python tools/gemm_table.py quick > gpurun_out/gemm_table.log 2>&1; echo "gemm_table rc=$?"
unset VT_EXPERIMENTAL
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench (shipping defaults) rc=$?"; tail -n 1 gpurun_out/bench.log | cut -c1-600
VT_TAIL_UNITS=1 VT_COLSUM_WIDE=1 VT_LN_BWD_V2=1 timeout 600 python bench.py --no-others > gpurun_out/bench_exp.log 2>&1; echo "bench (tail+colsum+ln2) rc=$?"; tail -n 1 gpurun_out/bench_exp.log | cut -c1-300
timeout 300 python tools/profile_step.py torchprof > gpurun_out/torchprof.log 2>&1; echo "torchprof rc=$?"
