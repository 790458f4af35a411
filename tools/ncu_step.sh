#!/bin/bash
# ncu evidence for one eager TimeSformer-B step (batch 8): launch list (durations of every kernel) + `--set full` captures of
# the hot kernels in the forward (first layers) and in the backward (last layers).  Run under gpurun, one GPU.
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r2_launches.csv python tools/profile_step.py ncu > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
PAT='regex:gemm|attn_tc|attn8|ln_bwd|ln_fwd|colsum|gelu|gather_cast'
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k "$PAT" -c 56 \
  -o gpurun_out/r2_fwd python tools/profile_step.py ncu > gpurun_out/ncu_fwd.log 2>&1; echo "full fwd rc=$?"
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k "$PAT" -s 190 -c 70 \
  -o gpurun_out/r2_bwd python tools/profile_step.py ncu > gpurun_out/ncu_bwd.log 2>&1; echo "full bwd rc=$?"
for f in fwd bwd; do
  ncu -i gpurun_out/r2_$f.ncu-rep --page raw --csv > gpurun_out/r2_${f}_raw.csv 2> /dev/null
  rm -f gpurun_out/r2_$f.ncu-rep          # ~100 MB each: gpurun only brings back 64 MiB, the raw page is what gets summarised
done
ls -la gpurun_out/ | tail -n 12
