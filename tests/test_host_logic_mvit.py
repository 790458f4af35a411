"""MaskFeat / MViT host logic (strided q/k/v slices, pooling bookkeeping, skip paths, gradient routing, state-dict
surface) on CPU against the reference-generated goldens, with the kernel table replaced by tests/emu_kernels.py."""
import pytest
import torch

from tests.conftest import check_grads, rel_err


def build(g):
    from videotransformer_pytorch_b200 import MaskFeat
    kw = dict(g.kwargs)
    for k in ('pool_q_stride_size', 'embed_dim_mul', 'atten_head_mul'):
        if k in kw:
            kw[k] = [list(r) for r in kw[k]]
    m = MaskFeat(**kw)
    m.load_state_dict(g.state(torch.float32), strict=True)
    return m


def test_block_plan_matches_reference_factory():
    from oracle import mvit_oracle as mo
    from videotransformer_pytorch_b200.maskfeat import mvit_block_plan
    for pq in (((1, 1, 2, 2), (3, 1, 2, 2)), ((1, 1, 2, 2), (3, 1, 2, 2), (14, 1, 2, 2))):
        cfg = mo.maskfeat_config(pool_q_stride_size=pq)          # pinned to the reference factory by make_golden.py
        plan = mvit_block_plan(16, 96, 1, [[1, 2.0], [3, 2.0], [14, 2.0]], [[1, 2.0], [3, 2.0], [14, 2.0]],
                               [list(r) for r in pq], [1, 8, 8], [3, 3, 3])
        for mine, ref in zip(plan, cfg['blocks']):
            assert (mine['dim'], mine['dim_out'], mine['heads'], mine['hidden']) == (ref['dim'], ref['dim_out'], ref['heads'], ref['hidden'])
            assert list(mine['stride_kv']) == ref['stride_kv']
            assert (list(mine['stride_q']) if mine['stride_q'] else []) == ref['stride_q']


@pytest.mark.parametrize('name', ['maskfeat_s32', 'maskfeat_s64'])
def test_state_dict_surface(maskfeat_golden, name):
    g = maskfeat_golden(name)
    m = build(g)
    ref = g.state(torch.float32)
    assert sorted(m.state_dict().keys()) == sorted(ref.keys())
    assert m.embed_dims == g.cfg['embed_dims'] and m.downsample_rate == g.cfg['downsample_rate']
    assert m.mvit.norm_embed.normalized_shape[0] == g.cfg['out_dim']
    assert m.no_weight_decay_keywords() == {'pos_embed', 'cls_token', 'mask_token'}


@pytest.mark.parametrize('name', ['maskfeat_s32', 'maskfeat_s64', 'maskfeat_s64_3stage'])
def test_forward_features(maskfeat_golden, emu, name):
    g = maskfeat_golden(name)
    m = build(g).train()
    with torch.no_grad():
        f = m.forward_features(g.x, g.mask)
        f0 = m.forward_features(g.x)
    assert f.shape == g.feats.shape
    assert rel_err(f, g.feats) < 5e-5
    assert rel_err(f0[:, 0], g.feats_nomask_cls) < 5e-5


@pytest.mark.parametrize('name', ['maskfeat_s32', 'maskfeat_s64'])
def test_forward_backward(maskfeat_golden, emu, name):
    g = maskfeat_golden(name)
    m = build(g).train()
    pred, loss = m(g.x, g.target.double(), g.mask, g.cube_marker)      # reference targets are fp64
    assert pred.shape == g.pred.shape
    assert rel_err(pred, g.pred) < 5e-5
    assert abs(loss.item() - g.loss) < 1e-4 * abs(g.loss)
    loss.backward()
    grads = {k: (p.grad if p.grad is not None else torch.zeros_like(p)) for k, p in m.named_parameters()}
    check_grads(grads, g, 2e-3)


@pytest.mark.parametrize('case', ['zero_mask', 'full_mask', 'batch3_shared_centres'])
def test_loss_edge_cases_vs_oracle(maskfeat_golden, emu, case):
    """Edge cases of MaskFeat.forward (video_transformer.py:889-901) against the oracle: nothing masked (loss 0 through the
    1e-5 guard), everything masked, an odd batch whose cubes share a centre frame."""
    from oracle import mvit_oracle as mo
    g = maskfeat_golden('maskfeat_s32')
    m = build(g).train()
    cfg = g.cfg
    t, h, w = cfg['thw'][0], cfg['thw'][1] // cfg['downsample_rate'], cfg['thw'][2] // cfg['downsample_rate']
    gen = torch.Generator().manual_seed(5)
    B = 3 if case == 'batch3_shared_centres' else 2
    x = torch.randn(B, cfg['num_frames'], 3, cfg['img_size'], cfg['img_size'], generator=gen)
    target = torch.randn(B, cfg['num_frames'], h, w, cfg['feature_dim'] // cfg['stride'][0], generator=gen)
    if case == 'zero_mask':
        mask = torch.zeros(B, t, h, w)
    elif case == 'full_mask':
        mask = torch.ones(B, t, h, w)
    else:
        mask = (torch.rand(B, t, h, w, generator=gen) < 0.5).float()
    markers = [[[0, 1], [0, 1]], [[t - 1, 1]], [[1, 2], [0, 3]]][:B]         # duplicated and overlapping cubes
    pred, loss = m(x, target, mask, markers)
    sd = {k: v.double() for k, v in g.state(torch.float32).items()}
    with torch.no_grad():
        pred_o, loss_o = mo.maskfeat_forward(sd, x.double(), target.double(), mask.double(), markers, cfg)
    assert rel_err(pred.detach(), pred_o) < 5e-5
    assert abs(float(loss) - float(loss_o)) < 1e-5 * max(1.0, abs(float(loss_o)))
    if case == 'zero_mask':
        assert float(loss) == 0.0
    loss.backward()                                    # gradients exist and are finite even when the loss is identically 0
    assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in m.parameters())


def test_fp64_targets_give_an_fp64_loss_like_the_reference(maskfeat_golden, emu):
    """dataset.py:190 builds the HOG targets as fp64 numpy arrays, so the reference's masked-MSE (video_transformer.py:899-901)
    is an fp64 tensor; with fp64 targets this package keeps the loss in fp64 too (fp32 targets -> fp32 loss)."""
    from videotransformer_pytorch_b200 import MaskFeat
    g = maskfeat_golden('maskfeat_s32')
    kw = dict(g.kwargs)
    for k in ('pool_q_stride_size', 'embed_dim_mul', 'atten_head_mul'):
        if k in kw:
            kw[k] = [list(r) for r in kw[k]]
    m = MaskFeat(**kw)
    m.load_state_dict(g.state(torch.float32), strict=True)
    m.train()
    _, l64 = m(g.x, g.target.double(), g.mask, g.cube_marker)
    _, l32 = m(g.x, g.target.float(), g.mask, g.cube_marker)
    assert l64.dtype == torch.float64 and l32.dtype == torch.float32
    assert abs(float(l64) - g.loss) < 2e-5 * abs(g.loss) and abs(float(l32) - g.loss) < 2e-5 * abs(g.loss)
    l64.backward()
    assert all(p.grad is not None for n, p in m.named_parameters())
