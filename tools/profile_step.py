"""Profile one TimeSformer-B fwd+bwd step (batch 8).
  python tools/profile_step.py torchprof   -> per-kernel device time table + CPU wall vs GPU busy
  ncu ... python tools/profile_step.py ncu -> 2 warm-up steps, then one step between cudaProfilerStart/Stop
"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import Trainee, IMG, NUM_CLASSES
T = 8

mode = sys.argv[1] if len(sys.argv) > 1 else 'torchprof'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device('cuda:0')
torch.manual_seed(0)
net = Trainee().to(dev).train()
x = torch.randn(B, T, 3, IMG, IMG, device=dev)
y = torch.randint(0, NUM_CLASSES, (B,), device=dev)


def step():
    for p in net.parameters():
        p.grad = None
    net(x, y).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
print('affinity cpus:', len(os.sched_getaffinity(0)), 'cpu_count:', os.cpu_count())
try:
    print('cgroup cpu.max:', open('/sys/fs/cgroup/cpu.max').read().strip())
except Exception as e:
    print('cgroup cpu.max: n/a', e)
if mode == 'ncu':
    # log every GEMM call of the profiled step (one kernel launch each, in order) so that tools/ncu_summarize.py can put the
    # algorithmic bytes of each shape next to the measured DRAM traffic
    import json
    from videotransformer_pytorch_b200 import _lib
    calls = []
    real = _lib.K.gemm

    def logged(a_, b_, M, N, Kd, **kw):
        calls.append(dict(M=M, N=N, K=Kd, a_mn=bool(kw.get('a_mn')), b_mn=bool(kw.get('b_mn')), epi=kw.get('epi', 'bf16'),
                          aux=kw.get('aux') is not None, split=bool(kw.get('split_ok')), tag=kw.get('tag')))
        return real(a_, b_, M, N, Kd, **kw)

    _lib.K.gemm = logged
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    _lib.K.gemm = real
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/r2_gemm_calls.json', 'w') as fh:
        json.dump(calls, fh)
else:
    t0 = time.perf_counter()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    print(f'wall per step: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms')
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    ev = [e for e in prof.key_averages() if e.device_time_total > 0]
    tot = sum(e.self_device_time_total for e in ev)
    print(f'GPU busy (sum of kernel time): {tot / 1e3:.2f} ms')
    rows = sorted(ev, key=lambda e: -e.self_device_time_total)[:40]
    for e in rows:
        print(f'{e.self_device_time_total / 1e3:9.3f} ms  x{e.count:4d}  {e.key[:110]}')
