"""2-GPU NCCL check of the data-parallel path: eager GradientBuckets and the CUDA-graph captured step both
produce the mean of the per-rank gradients.  -m gpu (skipped with fewer than 2 GPUs)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import datetime
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=120))
    from tests.test_gpu_graph import Net
    from videotransformer_pytorch_b200.ddp import GradientBuckets
    from videotransformer_pytorch_b200.graph import GraphedTrainStep
    torch.manual_seed(0)
    net = Net().to(dev).train()
    g = torch.Generator().manual_seed(50 + rank)
    x = torch.randn(2, 4, 3, 48, 48, generator=g).to(dev)
    y = torch.randint(0, 10, (2,), generator=g).to(dev)
    # local (unreduced) gradients of this rank, eager, DropPath seed shared by all ranks
    torch.manual_seed(7)
    net(x, y).backward()
    local = [p.grad.detach().clone() for p in net.parameters()]
    for p in net.parameters():
        p.grad = None
    mean = []
    for t in local:
        t = t.clone(); dist.all_reduce(t); mean.append(t / world)
    red = GradientBuckets(net, bucket_bytes=1 << 20)
    # eager bucketed all-reduce
    red.zero_grad()
    torch.manual_seed(7)
    net(x, y).backward()
    red.finish()
    torch.cuda.synchronize()
    e_eager = max(float((p.grad - m).abs().max() / (m.abs().max() + 1e-12)) for p, m in zip(net.parameters(), mean))
    # captured step
    step = GraphedTrainStep(net, (x, y), reducer=red)
    torch.manual_seed(7)
    step(x, y)
    torch.cuda.synchronize()
    e_graph = max(float((p.grad - m).abs().max() / (m.abs().max() + 1e-12)) for p, m in zip(net.parameters(), mean))
    q.put((rank, e_eager, e_graph))
    dist.barrier()
    step = None
    torch.cuda.synchronize()
    os._exit(0)


def test_two_gpu_gradient_mean_eager_and_graph():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, e_eager, e_graph in res:
        assert e_eager < 1e-5, (rank, e_eager)
        assert e_graph < 1e-5, (rank, e_graph)
