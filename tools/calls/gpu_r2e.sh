#!/bin/bash
# round-2 call E: rank-4 map fault under compute-sanitizer, second-generation pooling kernels, profiles, ncu, full bench
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
rm -f gpurun_out/*.ncu-rep
timeout 120 python tools/res_probe.py spatial 1 > gpurun_out/probe_sp.log 2>&1; echo "probe spatial rc=$?"; grep -v "^  File\|^    " gpurun_out/probe_sp.log | tail -n 8 | cut -c1-250
timeout 400 compute-sanitizer --tool memcheck --print-limit 3 python tools/res_probe.py spatial 1 > gpurun_out/sanitizer_sp.log 2>&1; echo "sanitizer rc=$?"
grep -A12 "=========" gpurun_out/sanitizer_sp.log | head -n 60 | cut -c1-220
timeout 900 python -m pytest tests/test_gpu_mvit.py tests/test_gpu_mvit_oracle.py -q -m gpu -k pool > gpurun_out/test_pool.log 2>&1; echo "pool tests rc=$?"; tail -n 6 gpurun_out/test_pool.log | cut -c1-250
for v in 0 1; do
  VT_POOL_V2=$v timeout 600 python tools/maskfeat_bench.py > gpurun_out/maskfeat_v$v.log 2>&1; echo "maskfeat_bench pool_v2=$v rc=$?"
  grep -E "ms/step|pool_|kernel time total" gpurun_out/maskfeat_v$v.log | cut -c1-200 | head -n 14
done
timeout 300 python tools/profile_step.py torchprof > gpurun_out/torchprof_base.log 2>&1; echo "torchprof base rc=$?"
VT_TMA_RES=1 VT_COLSUM_WIDE=1 timeout 300 python tools/profile_step.py torchprof > gpurun_out/torchprof_feat.log 2>&1; echo "torchprof feat rc=$?"
head -n 40 gpurun_out/torchprof_base.log | cut -c1-160
timeout 600 python tools/gemm_table.py quick > gpurun_out/gemm_table.log 2>&1; echo "gemm_table rc=$?"
VT_TMA_RES=1 VT_COLSUM_WIDE=1 timeout 900 python bench.py > gpurun_out/bench_feat.log 2>&1; echo "bench (RES + wide colsum) rc=$?"; tail -n 1 gpurun_out/bench_feat.log | cut -c1-2500
VT_TMA_RES=1 VT_COLSUM_WIDE=1 bash tools/ncu_step.sh
du -sh gpurun_out
