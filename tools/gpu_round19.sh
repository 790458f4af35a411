#!/bin/bash
set -u
mkdir -p gpurun_out
cd "${GRAFT_REPO_ROOT:-.}"
timeout 200 python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_ddp.py tests/test_gpu_graph.py -m gpu -q -x > gpurun_out/test_e.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/test_e.log | cut -c1-250
timeout 200 python -m pytest tests/test_gpu_modules.py -m gpu -q > gpurun_out/test_modules.log 2>&1; echo "test_modules rc=$?"; tail -n 2 gpurun_out/test_modules.log
timeout 200 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.log 2>&1; echo "bench n1 rc=$?"; tail -n 1 gpurun_out/bench_n1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['frac'])"
NCCL_DEBUG=WARN timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench n2 rc=$?"; tail -n 1 gpurun_out/bench_n2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])"
