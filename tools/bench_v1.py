#!/usr/bin/env python
"""FALLBACK copy of the benchmark as it ran on hardware early in round 2 (gpurun call A, 18.49 ms / step).  bench.py — which
adds the other BASELINE workloads and more roofline fields — executes this file in a fresh process if its own measurement
raises on one GPU, so that a headline line is always produced.  Original docstring:

Headline benchmark: clips/sec of the TimeSformer-B (divided space-time, 8x224x224) forward+backward hot
path on N B200 GPUs, next to the reference algorithm's CPU timing.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 8]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference [--steps K --warmup W]      # CPU arm (oracle port of the reference)

One JSON line on stdout (rank 0).  A "step" = one pass of the hot path over one synthetic batch: patch embed,
12 x (temporal attn, spatial attn, FFN), final norm, cls head + cross-entropy, full backward, and for N > 1
the bucketed gradient all-reduce.  The optimizer update is outside the metric (BASELINE.json: fwd+bwd).
  value : inputs resident in HBM before the timed region
  e2e   : same step through the public nn.Module API with the clip batch copied from pinned host memory and
          the loss read back to the host every step
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))     # this file lives in tools/
sys.path.insert(0, ROOT)

METRIC = 'clips/sec (BxTx3x224x224) TimeSformer-B fwd+bwd'
UNIT = 'clips/s'
T, IMG, NUM_CLASSES = 8, 224, 400
# algorithmic FLOPs per clip, fwd+bwd (MAC = 2 FLOP, bwd = 2x fwd)
FLOP_PER_CLIP = 1.175e12


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            p = json.load(fh)
        return dict(tflops=float(p['bf16_tflops_sustained']), burst=float(p['bf16_tflops']), hbm=float(p['hbm_gbs']),
                    source='measured (MEASURED_PEAKS.json, sustained bf16)')
    except Exception:
        return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, source='fallback (B200_PROFILING.md)')


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference TimeSformer (oracle/vt_oracle.py), all host threads
# ------------------------------------------------------------------------------------------------------
def cpu_step_factory(batch):
    from oracle import vt_oracle as O
    torch.manual_seed(0)
    cfg = dict(O.TIMESFORMER_B)
    sd = O.random_timesformer_state(cfg, seed=0)
    g = torch.Generator().manual_seed(1)
    head_w = (torch.randn(NUM_CLASSES, 768, generator=g) * 0.02).requires_grad_(True)
    head_b = torch.zeros(NUM_CLASSES, requires_grad=True)
    for v in sd.values():
        v.requires_grad_(True)
    x = torch.randn(batch, T, 3, IMG, IMG, generator=g)
    y = torch.randint(0, NUM_CLASSES, (batch,), generator=g)

    def step():
        for v in sd.values():
            v.grad = None
        feat = O.timesformer_forward(sd, x, cfg, training=True)
        loss = torch.nn.functional.cross_entropy(feat @ head_w.t() + head_b, y)
        loss.backward()
        return float(loss.detach())
    return step


def host_threads():
    """CPU threads this process can really use: min(affinity mask, cgroup v2 cpu.max quota)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    try:
        quota, period = open('/sys/fs/cgroup/cpu.max').read().split()
        if quota != 'max':
            n = max(1, min(n, int(int(quota) / int(period))))
    except Exception:
        pass
    return n


WORKLOAD = 'TimeSformer-B divided_space_time 8x224x224 fwd+bwd (+cls head, CE), train mode, DropPath 0..0.1'


def run_cpu(steps, warmup, batch=1):
    cores = host_threads()
    torch.set_num_threads(cores)
    step = cpu_step_factory(batch)
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    total = sum(times)
    return dict(value=batch * steps / total, ms_per_step=1e3 * total / steps, cores=cores,
                sample=f'{steps} timed step(s) of fwd+bwd on a {batch}-clip batch (fp32, torch CPU kernels, '
                       f'{cores} threads), {warmup} warm-up')


def main_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # bounded: each step is a 1-clip sample of the 8-clip workload
    steps = max(1, min(args.steps, 8))
    warm = max(1, min(args.warmup, 2))
    r = run_cpu(steps, warm, batch=1)
    line = {
        'metric': METRIC, 'value': r['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps, 'warmup': warm,
        'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
        'config': {'workload': WORKLOAD, 'batch_per_gpu': 1, 'global_batch': 1, 'parallelism': 'cpu',
                   'arm': 'oracle port of the reference TimeSformer on the host cores (fp32, torch CPU kernels); each step is a '
                          '1-clip sample of the 8-clip workload'},
        'cpu_baseline': {'value': r['value'], 'unit': UNIT, 'cores': r['cores'], 'kind': 'port', 'sample': r['sample']},
        'e2e': {'value': r['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), f'--query-gpu={q}', '--format=csv,noheader,nounits',
                                       '-lms', '100'], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.f.read().splitlines():
            c = [v.strip() for v in ln.split(',')]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1])); pw.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'power_w_max': max(pw), 'samples': len(sm),
                'reasons': sorted(reasons)}


class Trainee(torch.nn.Module):
    """TimeSformer-B + classification head, as built by the reference's VideoTransformer
    (model_trainer.py:53-82); loss = cross-entropy (training_step :204-206 without mixup)."""

    def __init__(self):
        super().__init__()
        from videotransformer_pytorch_b200 import ClassificationHead, TimeSformer
        self.model = TimeSformer(num_frames=T, img_size=IMG, patch_size=16, embed_dims=768, num_heads=12,
                                 num_transformer_layers=12, attention_type='divided_space_time')
        self.cls_head = ClassificationHead(NUM_CLASSES, 768, eval_metrics='finetune')
        with torch.no_grad():   # temporal_fc is zero-init in the reference: make the branch live
            for n, p in self.model.named_parameters():
                if 'temporal_fc' in n:
                    p.normal_(std=0.02)

    def forward(self, x, y):
        return self.cls_head.loss(self.model(x), y)          # skinny-GEMV head + fused softmax-CE kernels


def main_gpu(args):
    import torch.distributed as dist
    from videotransformer_pytorch_b200 import _lib
    from videotransformer_pytorch_b200.ddp import GradientBuckets

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the hot path has no CPU fallback; use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        import datetime
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))
    _lib.load_library()

    B = args.batch
    torch.manual_seed(0)
    net = Trainee().to(dev).train()
    reducer = GradientBuckets(net) if world > 1 else None
    if world > 1 and args.reserve_sms:
        _lib.set_reserved_sms(args.reserve_sms)      # room for the overlapped NCCL all-reduce kernels
    g = torch.Generator().manual_seed(100 + rank)
    x_host = torch.randn(B, T, 3, IMG, IMG, generator=g).pin_memory()
    y_host = torch.randint(0, NUM_CLASSES, (B,), generator=g).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    def zero():
        if reducer is not None:
            reducer.zero_grad()
        else:
            for p in net.parameters():
                p.grad = None

    def step(x, y):
        zero()
        loss = net(x, y)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    eager_step = step
    graphed = None
    if not args.no_graph:
        # whole step (fwd + bwd [+ bucket all-reduces]) captured once, replayed with one launch per step
        from videotransformer_pytorch_b200.graph import GraphedTrainStep
        graphed = GraphedTrainStep(net, (x_dev, y_dev), reducer=reducer, warmup=3)
        step = lambda x, y: graphed(x, y)
    for _ in range(max(args.warmup, 3)):
        step(x_dev, y_dev)
    barrier()

    sampler = ClockSampler(torch.cuda.current_device()) if rank == 0 else None
    l0 = _lib.launch_count()
    ms_dev = timed(lambda: step(x_dev, y_dev), args.steps)
    launches = (_lib.launch_count() - l0)
    if graphed is not None:      # replays launch the kernels recorded at capture time (the host-side counter is not touched)
        launches = graphed.kernels_per_replay * args.steps

    # End to end through the public API: every step's clip batch comes from pinned host memory and the loss goes back
    # to the host.  The copy of step i+1 is issued on a copy stream while step i computes (double-buffered device
    # staging), exactly what a DataLoader with pin_memory + non_blocking transfers gives the reference's training loop.
    copy_stream = torch.cuda.Stream(device=dev)
    xbuf = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    ybuf = [torch.empty_like(y_dev), torch.empty_like(y_dev)]
    arrived = [torch.cuda.Event(), torch.cuda.Event()]
    state = {'i': 0}

    def issue_copy(slot):
        # no wait needed: the slot's previous consumer (two steps ago) finished before that step's loss.item() returned
        with torch.cuda.stream(copy_stream):
            xbuf[slot].copy_(x_host, non_blocking=True)
            ybuf[slot].copy_(y_host, non_blocking=True)
            arrived[slot].record(copy_stream)

    def e2e_step():
        i = state['i']
        state['i'] = i + 1
        slot = i & 1
        torch.cuda.current_stream(dev).wait_event(arrived[slot])     # this step's input (host -> device) is here
        loss = step(xbuf[slot], ybuf[slot])
        issue_copy(slot ^ 1)                                         # next step's input travels during this step
        return float(loss.item())                                    # device -> host read of the loss

    issue_copy(0)
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if sampler else None

    # ---- roofline of the dominant kernel (gemm_tcgen05_kernel), measured live with CUDA events ----------
    roof = None
    if True:   # every rank runs the instrumented steps (they contain the bucket all-reduces); rank 0 reports
        pk = peaks()
        rec = []
        orig = _lib.K.gemm
        ext = {'external': True}

        def timed_gemm(a, b, M, N, Kd, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True, **ext), torch.cuda.Event(enable_timing=True, **ext)
            e0.record()
            out = orig(a, b, M, N, Kd, **kw)
            e1.record()
            rec.append((e0, e1, 2.0 * M * N * Kd))
            return out
        _lib.K.gemm = timed_gemm
        reps = 2
        timing = None
        try:
            probe = None
            if not args.no_graph:
                try:
                    # Preferred: the SAME step captured once more with an external CUDA-event record node before and
                    # after every GEMM launch on the capture stream; a replay yields the in-situ duration of each launch.
                    from videotransformer_pytorch_b200.graph import GraphedTrainStep
                    rec.clear()
                    probe = GraphedTrainStep(net, (x_dev, y_dev), reducer=reducer, warmup=0)
                except Exception as exc:          # e.g. external events unsupported by this torch build
                    sys.stderr.write(f'roofline: graph-event probe unavailable ({exc}); falling back to eager events\n')
                    probe = None
            if probe is not None:
                for _ in range(2):
                    probe(x_dev, y_dev)
                torch.cuda.synchronize()
                reps = 1
                timing = 'external CUDA-event nodes around every GEMM launch inside the replayed step graph'
            else:
                ext.clear()
                rec.clear()
                for _ in range(2):
                    # eager issue of the same step with every GEMM bracketed by CUDA events; a spin kernel keeps the
                    # GPU busy while the host queues the step
                    torch.cuda._sleep(120_000_000)
                    eager_step(x_dev, y_dev)
                torch.cuda.synchronize()
                timing = 'CUDA events around every GEMM of an eagerly issued step'
        finally:
            _lib.K.gemm = orig
        t_ms = sum(a.elapsed_time(b) for a, b, _ in rec)
        fl = sum(f for _, _, f in rec)
        ach = fl / (t_ms * 1e-3) / 1e12
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'gemm_traffic.json')) as fh:
                traffic = json.load(fh).get('dram_bytes_per_launch')
        except Exception:
            pass
        roof = {'kernel': 'gemm_tcgen05_kernel', 'bound': 'tensor', 'achieved': ach, 'peak': pk['tflops'],
                'unit': 'TFLOP/s', 'frac': ach / pk['tflops'], 'traffic': traffic,
                'launches_timed': len(rec), 'gemm_ms_per_step': t_ms / reps, 'gemm_flop_per_step': fl / reps, 'timing': timing,
                'peak_source': pk['source'],
                'whole_step_frac_of_tensor_roofline': (FLOP_PER_CLIP * B / (ms_dev / args.steps * 1e-3) / 1e12) / pk['tflops']}

    if world > 1:
        dist.barrier()
    cpu = None
    if rank == 0:
        cpu = run_cpu(steps=2, warmup=1, batch=1)
    if rank == 0:
        value = world * B * args.steps / (ms_dev * 1e-3)
        e2e = world * B * args.steps / (ms_e2e * 1e-3)
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_dev / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'batch_per_gpu': B, 'global_batch': B * world,
                       'parallelism': f'dp{world}', 'residual_stream': 'fp32', 'gemm_operands': 'bf16/fp32-accum',
                       'optimizer': 'excluded (metric is fwd+bwd)', 'launch': 'eager' if args.no_graph else 'cuda-graph replay (fwd+bwd captured once)', 'grad_allreduce': f'fp32 buckets, NCCL AVG, overlapped with backward inside the graph, {args.reserve_sms} SMs reserved' if world > 1 else 'n/a',
                       'l2': 'per-step working set ~5 GB >> 126 MB L2 (no flush needed)'},
            'e2e': {'value': e2e, 'unit': UNIT, 'ms_per_step': ms_e2e / args.steps,
                    'h2d_bytes_per_step': x_host.numel() * 4 + y_host.numel() * 8, 'd2h_bytes_per_step': 4},
            'gpu_launches': launches, 'clocks': clocks, 'roofline': roof,
            'cpu_baseline': {'value': cpu['value'], 'unit': UNIT, 'cores': cpu['cores'], 'kind': 'port', 'sample': cpu['sample']},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        # captured graphs hold NCCL kernels: release them before the communicator goes away
        step = eager_step = graphed = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=8, help='clips per GPU (BASELINE config 2: 8)')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--reserve-sms', type=int, default=0, help='SMs kept free of persistent GEMM CTAs when N > 1 (NCCL overlap)')
    ap.add_argument('--no-graph', action='store_true', help='issue the step kernel by kernel instead of replaying a CUDA graph')
    args = ap.parse_args()
    if args.impl == 'reference':
        return main_reference(args)
    return main_gpu(args)


if __name__ == '__main__':
    main()
