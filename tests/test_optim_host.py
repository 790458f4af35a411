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


@pytest.mark.parametrize('kind', ['sgd', 'adamw'])
def test_state_dict_round_trip_resumes_identically(emu, kind):
    """Momentum / Adam moments and the step count live in torch.optim.Optimizer.state: a checkpointed and restored
    optimizer continues exactly like the uninterrupted one (Lightning resume)."""
    import copy
    from videotransformer_pytorch_b200.optim import FusedAdamW, FusedSGD
    mk = (lambda ps: FusedSGD(ps, lr=0.05, momentum=0.9, nesterov=True, weight_decay=0.01)) if kind == 'sgd' else \
         (lambda ps: FusedAdamW(ps, lr=1e-2, weight_decay=0.05))
    a_p, b_p = make_params(3), make_params(3)
    a, b = mk(a_p), mk(b_p)

    def grads(step, ps):
        gg = torch.Generator().manual_seed(200 + step)
        for p in ps:
            p.grad = torch.randn(p.shape, generator=gg)

    for step in range(3):
        grads(step, a_p); grads(step, b_p)
        a.step(clip_grad=1.0); b.step(clip_grad=1.0)
    sd = copy.deepcopy(b.state_dict())
    assert len(sd['state']) == len(b_p)
    names = ('momentum_buffer',) if kind == 'sgd' else ('exp_avg', 'exp_avg_sq')
    assert all(n in sd['state'][0] for n in names) and float(sd['state'][0]['step']) == 3.0
    c_p = [torch.nn.Parameter(p.detach().clone()) for p in b_p]
    c = mk(c_p)
    c.load_state_dict(sd)
    for step in range(3, 6):
        grads(step, a_p); grads(step, c_p)
        a.step(clip_grad=1.0); c.step(clip_grad=1.0)
    for p, q in zip(a_p, c_p):
        assert torch.equal(p.detach(), q.detach())
    # a torch.optim.SGD checkpoint (momentum buffers, no step counter) is accepted: not treated as a first step
    if kind == 'sgd':
        t_p = make_params(4)
        t = torch.optim.SGD(t_p, lr=0.05, momentum=0.9, nesterov=True, weight_decay=0.01)
        f_p = make_params(4)
        for step in range(2):
            grads(step, t_p)
            t.step()
        f = mk([torch.nn.Parameter(p.detach().clone()) for p in t_p])
        f.load_state_dict(copy.deepcopy(t.state_dict()))
        f_p = [p for g in f.param_groups for p in g['params']]
        grads(5, t_p); grads(5, f_p)
        t.step(); f.step()
        for p, q in zip(t_p, f_p):
            assert rel_err(q.detach(), p.detach()) < 1e-6


def test_step_bumps_parameter_versions(emu):
    """The CUDA kernels write parameters through raw pointers; step() must still advance `param._version` so the bf16
    weight shadows (transformer.ShadowWeights, keyed on the version) are re-cast in eager training."""
    from videotransformer_pytorch_b200.optim import FusedSGD
    from videotransformer_pytorch_b200.transformer import ShadowWeights

    class RawWrites(type(emu)):
        """emulation that, like the real kernels, does not touch the version counters"""
        def opt_sgd(self, tbl, clip, momentum, nesterov, first_step):
            params = tbl['_params']
            saved = [p._version for p in params]
            for p in params:
                p.data.add_(-0.1)          # .data writes do not bump the version
            assert [p._version for p in params] == saved

        def cast_bf16(self, x):
            return x.bfloat16()

    from videotransformer_pytorch_b200 import _lib
    old = _lib.K
    _lib.K = RawWrites(exact=True)
    try:
        p = torch.nn.Parameter(torch.randn(8, 4))
        sh = ShadowWeights()
        w0 = sh.get('w', p)
        opt = FusedSGD([p], lr=0.1)
        p.grad = torch.ones_like(p)
        v0 = p._version
        opt.step()
        assert p._version > v0
        w1 = sh.get('w', p)
        assert not torch.equal(w0, w1) and torch.equal(w1, p.detach().bfloat16())
    finally:
        _lib.K = old
