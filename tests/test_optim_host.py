"""Fused clip + optimizer host logic on the CPU emulation: equals the reference flow
(clip_gradients, model_trainer.py:155-170, followed by torch.optim SGD-nesterov / AdamW, optimizer.py:33-38)."""
import pytest
import torch

from tests.conftest import rel_err


def reference_clip(params, clip_grad):
    norms = []
    for p in params:
        n = torch.norm(p.grad.detach(), 2)
        norms.append(n)
        coef = clip_grad / (n + 1e-6)
        if coef < 1:
            p.grad.data.mul_(coef)
    return torch.norm(torch.stack(norms), 2)


def make_params(seed, dtype=torch.float32, device='cpu'):
    g = torch.Generator().manual_seed(seed)
    shapes = [(70000,), (33, 17), (5,), (128, 96), (1, 1, 96)]
    return [torch.nn.Parameter(torch.randn(s, generator=g).to(dtype).to(device)) for s in shapes]


@pytest.mark.parametrize('kind', ['sgd', 'adamw'])
def test_fused_step_matches_reference_flow(emu, kind):
    from videotransformer_pytorch_b200.optim import FusedAdamW, FusedSGD
    ref_p, my_p = make_params(0), make_params(0)
    groups = lambda ps: [{'params': [ps[0], ps[2]], 'weight_decay': 0.0}, {'params': [ps[1], ps[3], ps[4]]}]
    if kind == 'sgd':
        ref = torch.optim.SGD(groups(ref_p), lr=0.05, momentum=0.9, nesterov=True, weight_decay=0.05)
        mine = FusedSGD(groups(my_p), lr=0.05, momentum=0.9, nesterov=True, weight_decay=0.05)
    else:
        ref = torch.optim.AdamW(groups(ref_p), lr=1e-2, betas=(0.9, 0.999), weight_decay=0.05)
        mine = FusedAdamW(groups(my_p), lr=1e-2, betas=(0.9, 0.999), weight_decay=0.05)
    order = [0, 2, 1, 3, 4]                         # group order == the reference's clip order for these groups
    for step in range(4):
        gg = torch.Generator().manual_seed(100 + step)
        for a, b in zip(ref_p, my_p):
            gr = torch.randn(a.shape, generator=gg) * (3.0 if step % 2 else 0.05)     # some clipped, some not
            a.grad, b.grad = gr.clone(), gr.clone()
        if step == 2:                               # schedulers rewrite the groups in place (model_trainer.py:150-153)
            for o in (ref, mine):
                o.param_groups[1]['weight_decay'] = 0.01
                o.param_groups[0]['lr'] = o.param_groups[1]['lr'] = 0.02
        total_ref = reference_clip([ref_p[i] for i in order], 0.7)
        ref.step()
        total = mine.step(clip_grad=0.7)
        assert abs(float(total) - float(total_ref)) < 1e-5 * float(total_ref)
        for a, b in zip(ref_p, my_p):
            assert rel_err(b.detach(), a.detach()) < 2e-6, (kind, step)


def test_step_without_clipping_and_missing_grad(emu):
    from videotransformer_pytorch_b200.optim import FusedSGD
    ps, qs = make_params(1), make_params(1)
    ref = torch.optim.SGD(ps, lr=0.1, momentum=0.9, nesterov=True, weight_decay=0.0)
    mine = FusedSGD(qs, lr=0.1)
    for a, b in zip(ps, qs):
        a.grad = torch.ones_like(a)
        b.grad = torch.ones_like(b)
    ref.step()
    assert mine.step() is None
    for a, b in zip(ps, qs):
        assert rel_err(b.detach(), a.detach()) < 1e-6
    qs[0].grad = None
    with pytest.raises(RuntimeError, match='no gradient'):
        mine.step()
