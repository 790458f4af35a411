"""MaskFeat / MViT oracle (oracle/mvit_oracle.py): against the reference-generated goldens, and — for the restated
pytorchvideo block arithmetic, which the reference does not vendor — against torchvision's independent MViT."""
import pytest
import torch

from tests.conftest import check_grads, rel_err


@pytest.mark.parametrize('name', ['maskfeat_s32', 'maskfeat_s64', 'maskfeat_s64_3stage'])
def test_maskfeat_oracle_vs_golden(maskfeat_golden, name):
    from oracle import mvit_oracle as mo
    g = maskfeat_golden(name)
    sd = {k: v.requires_grad_(True) for k, v in g.state().items()}
    x, mask, target = g.x.double(), g.mask.double(), g.target.double()
    feats = mo.maskfeat_forward_features(sd, x, mask, g.cfg)
    assert rel_err(feats, g.feats) < 1e-12
    assert rel_err(mo.maskfeat_forward_features(sd, x, None, g.cfg)[:, 0], g.feats_nomask_cls) < 1e-12
    pred, loss = mo.maskfeat_forward(sd, x, target, mask, g.cube_marker, g.cfg)
    assert rel_err(pred, g.pred) < 1e-12
    assert abs(loss.item() - g.loss) < 1e-12 * max(1.0, abs(g.loss))
    if g.grad or g.gradsum:
        loss.backward()
        check_grads({k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in sd.items()}, g, 1e-5)


def test_reference_config_shapes():
    """The model the reference trains (model_trainer.py:54): 16 blocks, dims 96->768, head dim 96 throughout,
    (8,56,56) -> (8,14,14) tokens, K/V pooled to 392-1568 tokens."""
    from oracle import mvit_oracle as mo
    cfg = mo.maskfeat_config()
    assert cfg['thw'] == (8, 56, 56) and cfg['downsample_rate'] == 4 and cfg['embed_dims'] == 768
    assert [b['dim'] for b in cfg['blocks']] == [96, 192, 192] + [384] * 11 + [768] * 2
    assert [b['dim_out'] for b in cfg['blocks']] == [192, 192, 384] + [384] * 10 + [768] * 3
    assert all(b['dim'] // b['heads'] == 96 for b in cfg['blocks'])
    assert [b['stride_q'] for b in cfg['blocks']] == [[], [1, 2, 2], [], [1, 2, 2]] + [[]] * 12
    assert [b['stride_kv'] for b in cfg['blocks']] == [[1, 8, 8]] + [[1, 4, 4]] * 2 + [[1, 2, 2]] * 13


def _torchvision_mvit(cfg):
    from torchvision.models.video.mvit import MSBlockConfig, MViT
    from functools import partial
    settings = [MSBlockConfig(num_heads=b['heads'], input_channels=b['dim'], output_channels=b['dim_out'],
                              kernel_q=list(b['kernel_q']), kernel_kv=list(b['kernel_kv']),
                              stride_q=list(b['stride_q']), stride_kv=list(b['stride_kv'])) for b in cfg['blocks']]
    return MViT(spatial_size=(cfg['img_size'], cfg['img_size']), temporal_size=cfg['num_frames'],
                block_setting=settings, residual_pool=False, residual_with_cls_embed=False, rel_pos_embed=False,
                proj_after_attn=False, dropout=0.0, num_classes=4,
                norm_layer=partial(torch.nn.LayerNorm, eps=cfg['block_norm_eps']),
                patch_embed_kernel=cfg['kernel'], patch_embed_stride=cfg['stride'], patch_embed_padding=cfg['padding'])


def _to_torchvision_keys(sd, cfg):
    out = {'conv_proj.weight': sd['patch_embed.patch_model.weight'], 'conv_proj.bias': sd['patch_embed.patch_model.bias'],
           'pos_encoding.class_token': sd['mvit.cls_positional_encoding.cls_token'].reshape(-1),
           'pos_encoding.spatial_pos': sd['mvit.cls_positional_encoding.pos_embed_spatial'][0],
           'pos_encoding.temporal_pos': sd['mvit.cls_positional_encoding.pos_embed_temporal'][0],
           'pos_encoding.class_pos': sd['mvit.cls_positional_encoding.pos_embed_class'].reshape(-1),
           'norm.weight': sd['mvit.norm_embed.weight'], 'norm.bias': sd['mvit.norm_embed.bias']}
    for i, b in enumerate(cfg['blocks']):
        s, d = f'mvit.blocks.{i}.', f'blocks.{i}.'
        for n in ('norm1', 'norm2'):
            out[d + n + '.weight'], out[d + n + '.bias'] = sd[s + n + '.weight'], sd[s + n + '.bias']
        out[d + 'attn.qkv.weight'] = torch.cat([sd[s + f'attn.{n}.weight'] for n in 'qkv'], 0)
        out[d + 'attn.qkv.bias'] = torch.cat([sd[s + f'attn.{n}.bias'] for n in 'qkv'], 0)
        out[d + 'attn.project.0.weight'], out[d + 'attn.project.0.bias'] = sd[s + 'attn.proj.weight'], sd[s + 'attn.proj.bias']
        for n in 'qkv':
            if s + f'attn.pool_{n}.weight' in sd:
                out[d + f'attn.pool_{n}.pool.weight'] = sd[s + f'attn.pool_{n}.weight']
                out[d + f'attn.pool_{n}.norm_act.0.weight'] = sd[s + f'attn.norm_{n}.weight']
                out[d + f'attn.pool_{n}.norm_act.0.bias'] = sd[s + f'attn.norm_{n}.bias']
        out[d + 'mlp.0.weight'], out[d + 'mlp.0.bias'] = sd[s + 'mlp.fc1.weight'], sd[s + 'mlp.fc1.bias']
        out[d + 'mlp.3.weight'], out[d + 'mlp.3.bias'] = sd[s + 'mlp.fc2.weight'], sd[s + 'mlp.fc2.bias']
        if s + 'proj.weight' in sd:
            out[d + 'project.weight'], out[d + 'project.bias'] = sd[s + 'proj.weight'], sd[s + 'proj.bias']
    return out


@pytest.mark.parametrize('three_stage', [False, True])
def test_restated_blocks_match_torchvision_mvit(three_stage):
    """Independent cross-check of the restated pytorchvideo arithmetic (a14): torchvision's MViT-v1 is a separate
    implementation of the same network; with identical weights (pool norms at eps 1e-6, torchvision's only option)
    the token features must agree to rounding."""
    from oracle import mvit_oracle as mo
    kw = dict(img_size=64, num_frames=4, pool_norm_eps=1e-6)
    if three_stage:
        kw['pool_q_stride_size'] = ((1, 1, 2, 2), (3, 1, 2, 2), (14, 1, 2, 2))
    cfg = mo.maskfeat_config(**kw)
    sd = mo.random_maskfeat_state(cfg, seed=11, dtype=torch.float64)
    tv = _torchvision_mvit(cfg).double().eval()
    mapped = _to_torchvision_keys(sd, cfg)
    mapped['head.1.weight'], mapped['head.1.bias'] = tv.head[1].weight.data, tv.head[1].bias.data
    tv.load_state_dict(mapped, strict=True)
    x = torch.randn(2, cfg['num_frames'], 3, 64, 64, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    captured = {}
    tv.norm.register_forward_hook(lambda m, i, o: captured.__setitem__('feats', o))
    with torch.no_grad():
        tv(x.transpose(1, 2))
        mine = mo.maskfeat_forward_features(sd, x, None, cfg)
    assert mine.shape == captured['feats'].shape
    assert rel_err(mine, captured['feats']) < 1e-10
