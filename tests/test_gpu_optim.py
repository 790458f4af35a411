"""Fused clip + optimizer kernels against the reference flow (clip_gradients + torch.optim) computed on CPU.  -m gpu"""
import pytest
import torch

from tests.conftest import rel_err
from tests.test_optim_host import make_params, reference_clip

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('kind', ['sgd', 'adamw'])
@pytest.mark.parametrize('clip', [0.7, None])
def test_fused_step_matches_reference_flow(kind, clip):
    from videotransformer_pytorch_b200.optim import FusedAdamW, FusedSGD
    ref_p, my_p = make_params(0), make_params(0, device='cuda')
    groups = lambda ps: [{'params': [ps[0], ps[2]], 'weight_decay': 0.0}, {'params': [ps[1], ps[3], ps[4]]}]
    if kind == 'sgd':
        ref = torch.optim.SGD(groups(ref_p), lr=0.05, momentum=0.9, nesterov=True, weight_decay=0.05)
        mine = FusedSGD(groups(my_p), lr=0.05, momentum=0.9, nesterov=True, weight_decay=0.05)
    else:
        ref = torch.optim.AdamW(groups(ref_p), lr=1e-2, betas=(0.9, 0.999), weight_decay=0.05)
        mine = FusedAdamW(groups(my_p), lr=1e-2, betas=(0.9, 0.999), weight_decay=0.05)
    order = [0, 2, 1, 3, 4]
    for step in range(4):
        gg = torch.Generator().manual_seed(100 + step)
        for a, b in zip(ref_p, my_p):
            gr = torch.randn(a.shape, generator=gg) * (3.0 if step % 2 else 0.05)
            a.grad, b.grad = gr.clone(), gr.cuda()
        if step == 2:
            for o in (ref, mine):
                o.param_groups[1]['weight_decay'] = 0.01
                o.param_groups[0]['lr'] = o.param_groups[1]['lr'] = 0.02
        if clip is not None:
            total_ref = reference_clip([ref_p[i] for i in order], clip)
        ref.step()
        total = mine.step(clip_grad=clip)
        if clip is not None:
            assert abs(float(total) - float(total_ref)) < 1e-5 * float(total_ref)
        else:
            assert total is None
        for a, b in zip(ref_p, my_p):
            assert rel_err(b.detach().cpu(), a.detach()) < 5e-6, (kind, step)


def test_fused_sgd_on_a_model_with_static_grads():
    """Unaligned tensor sizes, many tensors, gradients that keep their address between steps."""
    from videotransformer_pytorch_b200.optim import FusedSGD
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.LayerNorm(53), torch.nn.Linear(53, 7))
    ref = torch.nn.Sequential(torch.nn.Linear(37, 53), torch.nn.LayerNorm(53), torch.nn.Linear(53, 7))
    ref.load_state_dict(net.state_dict())
    net = net.cuda()
    ro = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4)
    mo = FusedSGD(net.parameters(), lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4)
    static = [torch.zeros_like(p) for p in net.parameters()]
    for p, g in zip(net.parameters(), static):
        p.grad = g
    for step in range(3):
        gen = torch.Generator().manual_seed(step)
        for p, q, g in zip(ref.parameters(), net.parameters(), static):
            gr = torch.randn(p.shape, generator=gen)
            p.grad = gr.clone()
            g.copy_(gr)
        reference_clip(list(ref.parameters()), 1.0)
        ro.step()
        mo.step(clip_grad=1.0)
    for p, q in zip(ref.parameters(), net.parameters()):
        assert rel_err(q.detach().cpu(), p.detach()) < 5e-6


def test_eager_training_with_fused_sgd_tracks_torch_sgd():
    """Eager (non-graph) training of a package model: the fused optimizer updates parameters through raw pointers, the
    bf16 weight shadows must follow (stale-shadow bug).  Two identical TimeSformers, one stepped by
    FusedSGD and one by torch.optim.SGD, stay together for 3 steps and both move away from the initial weights."""
    from videotransformer_pytorch_b200 import TimeSformer
    from videotransformer_pytorch_b200.optim import FusedSGD
    cfg = dict(num_frames=4, img_size=48, patch_size=16, embed_dims=128, num_heads=2, num_transformer_layers=2)
    torch.manual_seed(0)
    a = TimeSformer(**cfg)
    with torch.no_grad():
        for n, p in a.named_parameters():
            if 'temporal_fc' in n:
                p.normal_(std=0.05)
    b = TimeSformer(**cfg)
    b.load_state_dict(a.state_dict())
    a, b = a.cuda().eval(), b.cuda().eval()          # eval: no DropPath randomness between the two copies
    oa = FusedSGD(a.parameters(), lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-4)
    ob = torch.optim.SGD(b.parameters(), lr=0.05, momentum=0.9, nesterov=True, weight_decay=1e-4)
    x = torch.randn(2, 4, 3, 48, 48, generator=torch.Generator().manual_seed(1)).cuda()
    losses = []
    for step in range(3):
        for m, o in ((a, oa), (b, ob)):
            for p in m.parameters():
                p.grad = None
            loss = m(x).square().mean()
            loss.backward()
            o.step()
            losses.append(float(loss))
    la, lb = losses[0::2], losses[1::2]
    assert abs(la[0] - lb[0]) < 1e-6 * abs(lb[0])
    assert la[2] != la[0]                                             # the parameters really moved
    for i in range(3):
        assert abs(la[i] - lb[i]) < 2e-3 * abs(lb[i]), (la, lb)      # stale bf16 shadows would freeze la at la[0]
    for (n, p), q in zip(a.named_parameters(), b.parameters()):
        assert rel_err(p.detach(), q.detach()) < 1e-3, n
