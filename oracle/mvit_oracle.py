"""Plain-torch CPU restatement of the reference's MaskFeat / MViT path.

TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY STATUS
* a13 (conv patch embed + mask-token mixing, video_transformer.py:911-922, :585-618) and a15 (decoder + masked
  MSE loss, :876-909) live in the reference tree: ``oracle/make_golden.py`` runs the REAL ``MaskFeat`` class and asserts
  this file agrees with it to 1e-12 (fp64).
* a14 — the block arithmetic — lives in the third-party dependency **pytorchvideo** (unpinned in the reference's
  requirements; README.md:166 points at tree 9d0ca900f0427ed9b47b6182ad05f75c0e66274b; the keyword set passed at
  video_transformer.py:764-785 matches the 0.1.3-era ``MultiScaleBlock``), which is absent from /root/reference and
  from this image.  Its published algorithm (``pytorchvideo/layers/attention.py``: ``Mlp``, ``_attention_pool``,
  ``MultiScaleAttention``, ``MultiScaleBlock``; ``layers/positional_encoding.py``:
  ``SpatioTemporalClsPositionalEncoding``; ``layers/utils.py``: ``round_width``; ``models/vision_transformers.py``:
  ``MultiscaleVisionTransformers.forward``) is restated here.  There are no reference tests or vectors for it:
  **parity unpinned** for a14.  As an independent cross-check ``tests/test_oracle_mvit.py`` loads the same weights into
  torchvision's ``MViT`` (v1 settings), a separate implementation of the same published network, and requires agreement
  to 1e-10 — that checks the structure, it is not the reference.

Conventions: functional, consumes a reference-format state dict (keys of ``MaskFeat.state_dict()``,
video_transformer.py:834-857), torch primitives only; backward is autograd.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .vt_oracle import gelu_erf, layer_norm, linear

Tensor = torch.Tensor


# ----------------------------------------------------------------------------
# configuration (restates create_multiscale_vision_transformers, video_transformer.py:693-800)
# ----------------------------------------------------------------------------
def round_width(width, multiplier, min_width=8, divisor=8):
    """pytorchvideo.layers.utils.round_width (ceil=False branch)."""
    if not multiplier:
        return int(width)
    width = float(width) * float(multiplier)
    min_width = min_width or divisor
    out = max(min_width, int(width + divisor / 2) // divisor * divisor)
    if out < 0.9 * width:
        out += divisor
    return int(out)


def maskfeat_config(img_size=224, num_frames=16, feature_dim=216, patch_embed_dim=96, depth=16, num_heads=1,
                    conv_patch_embed_kernel=(3, 7, 7), conv_patch_embed_stride=(2, 4, 4),
                    conv_patch_embed_padding=(1, 3, 3),
                    embed_dim_mul=((1, 2.0), (3, 2.0), (14, 2.0)), atten_head_mul=((1, 2.0), (3, 2.0), (14, 2.0)),
                    pool_q_stride_size=((1, 1, 2, 2), (3, 1, 2, 2)), pool_kv_stride_adaptive=(1, 8, 8),
                    pool_kvq_kernel=(3, 3, 3), mlp_ratio=4.0, pool_norm_eps=1e-5, block_norm_eps=1e-6) -> dict:
    """Defaults = the model built at model_trainer.py:54 (``pool_q_stride_size`` overridden to two stages,
    ``feature_dim=2*2*2*3*9``) on top of MaskFeat.__init__ defaults (video_transformer.py:810-826).

    ``pool_norm_eps``: the reference leaves ``attn_norm_layer`` commented out (video_transformer.py:773), so the
    q/k/v pooling norms are plain ``nn.LayerNorm`` (eps 1e-5); block norms and the final norm use eps 1e-6 (:660-662).
    """
    dim_mul = [1.0] * (depth + 1)
    head_mul = [1.0] * (depth + 1)
    for i, m in embed_dim_mul:                                   # :707-713
        dim_mul[int(i)] = float(m)
    for i, m in atten_head_mul:
        head_mul[int(i)] = float(m)
    stride_q: List[List[int]] = [[] for _ in range(depth)]
    kernel_q: List[List[int]] = [[] for _ in range(depth)]
    for row in pool_q_stride_size:                               # :722-730
        stride_q[int(row[0])] = [int(s) for s in row[1:]]
        kernel_q[int(row[0])] = list(pool_kvq_kernel) if pool_kvq_kernel is not None else \
            [s + 1 if s > 1 else s for s in row[1:]]
    stride_kv: List[List[int]] = [[] for _ in range(depth)]
    kernel_kv: List[List[int]] = [[] for _ in range(depth)]
    cur = list(pool_kv_stride_adaptive)                          # :733-742
    for i in range(depth):
        if len(stride_q[i]) > 0:
            cur = [max(cur[d] // stride_q[i][d], 1) for d in range(3)]
        stride_kv[i] = list(cur)
        kernel_kv[i] = list(pool_kvq_kernel) if pool_kvq_kernel is not None else [s + 1 if s > 1 else s for s in cur]
    blocks = []
    heads, dim = num_heads, patch_embed_dim
    for i in range(depth):                                       # :754-761
        heads = round_width(heads, head_mul[i], min_width=1, divisor=1)
        dim = round_width(dim, dim_mul[i], divisor=heads)
        dim_out = round_width(dim, dim_mul[i + 1], divisor=round_width(heads, head_mul[i + 1]))
        blocks.append(dict(dim=dim, dim_out=dim_out, heads=heads, kernel_q=kernel_q[i], stride_q=stride_q[i],
                           kernel_kv=kernel_kv[i], stride_kv=stride_kv[i], hidden=int(dim * mlp_ratio)))
    st, sh, sw = conv_patch_embed_stride
    return dict(img_size=img_size, num_frames=num_frames, feature_dim=feature_dim, patch_embed_dim=patch_embed_dim,
                depth=depth, kernel=tuple(conv_patch_embed_kernel), stride=tuple(conv_patch_embed_stride),
                padding=tuple(conv_patch_embed_padding), blocks=blocks,
                thw=(num_frames // st, img_size // sh, img_size // sw),          # :690-692
                downsample_rate=2 ** len(pool_q_stride_size),                     # :829
                embed_dims=2 ** len(embed_dim_mul) * patch_embed_dim,             # :830
                out_dim=blocks[-1]['dim_out'], pool_norm_eps=pool_norm_eps, block_norm_eps=block_norm_eps)


# ----------------------------------------------------------------------------
# pytorchvideo blocks (a14, restated)
# ----------------------------------------------------------------------------
def attention_pool(t: Tensor, thw: Sequence[int], *, conv_w: Optional[Tensor] = None, stride=None,
                   max_kernel=None, norm_w: Optional[Tensor] = None, norm_b: Optional[Tensor] = None,
                   eps: float = 1e-5) -> Tuple[Tensor, Tuple[int, int, int]]:
    """pytorchvideo ``_attention_pool`` with ``has_cls_embed=True``.

    ``t``: [B, heads, 1+T*H*W, C] (or [B, 1+L, C], treated as one head).  The cls row bypasses the pool; the norm
    (when given) is applied to every row including cls.  Either a depthwise ``conv_w`` [C,1,kt,kh,kw] (padding k//2)
    or a max pool of ``max_kernel`` (padding k//2) with ``stride``.
    """
    three = t.ndim == 3
    if three:
        t = t.unsqueeze(1)
    cls, body = t[:, :, :1, :], t[:, :, 1:, :]
    B, Hh, L, C = body.shape
    T, H, W = thw
    assert L == T * H * W
    vol = body.reshape(B * Hh, T, H, W, C).permute(0, 4, 1, 2, 3)
    if conv_w is not None:
        pad = [k // 2 for k in conv_w.shape[2:]]
        vol = F.conv3d(vol, conv_w, None, stride=tuple(stride), padding=tuple(pad), groups=C)
    else:
        pad = [k // 2 for k in max_kernel]
        vol = F.max_pool3d(vol, tuple(max_kernel), tuple(stride), tuple(pad))
    new_thw = (vol.shape[2], vol.shape[3], vol.shape[4])
    body = vol.reshape(B, Hh, C, -1).transpose(2, 3)
    out = torch.cat([cls, body], dim=2)
    if norm_w is not None:
        out = layer_norm(out, norm_w, norm_b, eps)
    if three:
        out = out.squeeze(1)
    return out, new_thw


def multiscale_attention(sd: Dict[str, Tensor], pre: str, x: Tensor, thw, blk: dict, eps_pool: float):
    """pytorchvideo ``MultiScaleAttention.forward`` (pool_first=False, separate q/k/v Linear, conv pooling)."""
    B, N, C = x.shape
    Hh = blk['heads']
    hd = C // Hh

    def heads(name):
        y = linear(x, sd[pre + name + '.weight'], sd[pre + name + '.bias'])
        return y.reshape(B, N, Hh, hd).permute(0, 2, 1, 3)

    q, k, v = heads('q'), heads('k'), heads('v')
    q_thw = tuple(thw)
    if len(blk['stride_q']) > 0 and (math.prod(blk['kernel_q']) > 1 or math.prod(blk['stride_q']) > 1):
        q, q_thw = attention_pool(q, thw, conv_w=sd[pre + 'pool_q.weight'], stride=blk['stride_q'],
                                  norm_w=sd[pre + 'norm_q.weight'], norm_b=sd[pre + 'norm_q.bias'], eps=eps_pool)
    if len(blk['stride_kv']) > 0 and (math.prod(blk['kernel_kv']) > 1 or math.prod(blk['stride_kv']) > 1):
        k, _ = attention_pool(k, thw, conv_w=sd[pre + 'pool_k.weight'], stride=blk['stride_kv'],
                              norm_w=sd[pre + 'norm_k.weight'], norm_b=sd[pre + 'norm_k.bias'], eps=eps_pool)
        v, _ = attention_pool(v, thw, conv_w=sd[pre + 'pool_v.weight'], stride=blk['stride_kv'],
                              norm_w=sd[pre + 'norm_v.weight'], norm_b=sd[pre + 'norm_v.bias'], eps=eps_pool)
    attn = (q @ k.transpose(-2, -1)) * (hd ** -0.5)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B, q.shape[2], C)
    out = linear(out, sd[pre + 'proj.weight'], sd[pre + 'proj.bias'])
    return out, q_thw


def multiscale_block(sd: Dict[str, Tensor], pre: str, x: Tensor, thw, blk: dict, eps_block: float, eps_pool: float):
    """pytorchvideo ``MultiScaleBlock.forward`` (drop-path rate 0 in the reference: video_transformer.py:642)."""
    xn = layer_norm(x, sd[pre + 'norm1.weight'], sd[pre + 'norm1.bias'], eps_block)
    x_block, new_thw = multiscale_attention(sd, pre + 'attn.', xn, thw, blk, eps_pool)
    if len(blk['stride_q']) > 0:
        kernel_skip = [s + 1 if s > 1 else s for s in blk['stride_q']]
        x_res, _ = attention_pool(x, thw, max_kernel=kernel_skip, stride=blk['stride_q'])
    else:
        x_res = x
    x = x_res + x_block
    xn = layer_norm(x, sd[pre + 'norm2.weight'], sd[pre + 'norm2.bias'], eps_block)
    h = gelu_erf(linear(xn, sd[pre + 'mlp.fc1.weight'], sd[pre + 'mlp.fc1.bias']))
    x_mlp = linear(h, sd[pre + 'mlp.fc2.weight'], sd[pre + 'mlp.fc2.bias'])
    if blk['dim'] != blk['dim_out']:
        x = linear(xn, sd[pre + 'proj.weight'], sd[pre + 'proj.bias'])
    return x + x_mlp, new_thw


def cls_positional_encoding(sd: Dict[str, Tensor], pre: str, x: Tensor, thw) -> Tensor:
    """pytorchvideo ``SpatioTemporalClsPositionalEncoding.forward`` (sep_pos_embed=True, has_cls=True)."""
    B = x.shape[0]
    T, H, W = thw
    x = torch.cat([sd[pre + 'cls_token'].expand(B, -1, -1), x], dim=1)
    pos = sd[pre + 'pos_embed_spatial'].repeat(1, T, 1) + \
        torch.repeat_interleave(sd[pre + 'pos_embed_temporal'], H * W, dim=1)
    pos = torch.cat([sd[pre + 'pos_embed_class'], pos], dim=1)
    return x + pos


def mvit_forward(sd: Dict[str, Tensor], x: Tensor, cfg: dict, pre: str = 'mvit.', return_blocks: bool = False):
    """pytorchvideo ``MultiscaleVisionTransformers.forward`` with ``patch_embed=None`` (video_transformer.py:681,:793),
    no pos-drop, no patch-embed norm, no head."""
    x = cls_positional_encoding(sd, pre + 'cls_positional_encoding.', x, cfg['thw'])
    thw = tuple(cfg['thw'])
    trace = []
    for i, blk in enumerate(cfg['blocks']):
        x, thw = multiscale_block(sd, f'{pre}blocks.{i}.', x, thw, blk, cfg['block_norm_eps'], cfg['pool_norm_eps'])
        if return_blocks:
            trace.append(x)
    x = layer_norm(x, sd[pre + 'norm_embed.weight'], sd[pre + 'norm_embed.bias'], cfg['block_norm_eps'])
    return (x, trace) if return_blocks else x


# ----------------------------------------------------------------------------
# MaskFeat (a13, a15 — in-tree reference code, pinned by make_golden.py)
# ----------------------------------------------------------------------------
def conv_patch_embed(sd: Dict[str, Tensor], x: Tensor, cfg: dict) -> Tensor:
    """PatchEmbeding.forward, video_transformer.py:578-581, on ``x.transpose(1,2)`` (:912).  x: [B,T,3,H,W]."""
    y = F.conv3d(x.transpose(1, 2), sd['patch_embed.patch_model.weight'], sd['patch_embed.patch_model.bias'],
                 stride=cfg['stride'], padding=cfg['padding'])
    return y.flatten(2).transpose(1, 2)


def maskfeat_forward_features(sd: Dict[str, Tensor], x: Tensor, mask: Optional[Tensor], cfg: dict) -> Tensor:
    """MaskFeat.forward_features, video_transformer.py:911-922."""
    t = conv_patch_embed(sd, x, cfg)
    if mask is not None:
        r = cfg['downsample_rate']
        dense = mask.repeat_interleave(r, dim=2).repeat_interleave(r, dim=3)       # nearest-neighbour up-sampling :917
        w = dense.flatten(1).unsqueeze(-1).to(t.dtype)
        t = t * (1 - w) + sd['mask_token'] * w                                     # :919
    return mvit_forward(sd, t, cfg)


def center_frame_mask(mask: Tensor, cube_marker, stride_t: int, num_frames: int) -> Tensor:
    """video_transformer.py:889-896: repeat the mask over dt and keep only each cube's centre frame."""
    m = mask.repeat_interleave(stride_t, dim=1).clone()
    for i, cubes in enumerate(cube_marker):
        keep = torch.zeros(num_frames, dtype=torch.bool)
        for start, span in cubes:
            keep[start * stride_t + span * stride_t // 2] = True
        m[i, ~keep] = 0
    return m


def maskfeat_forward(sd: Dict[str, Tensor], x: Tensor, target: Tensor, mask: Tensor, cube_marker, cfg: dict):
    """MaskFeat.forward, video_transformer.py:876-909 (visualize=False).  Returns (pred [B,T,h,w,dc], loss)."""
    f = maskfeat_forward_features(sd, x, mask, cfg)
    p = linear(f, sd['decoder_pred.weight'], sd['decoder_pred.bias'])[:, 1:, :]      # :878-879
    B = p.shape[0]
    dt = cfg['stride'][0]
    t = cfg['num_frames'] // dt
    h = cfg['img_size'] // (cfg['stride'][1] * cfg['downsample_rate'])
    w = cfg['img_size'] // (cfg['stride'][2] * cfg['downsample_rate'])
    dc = p.shape[-1] // dt
    p = p.reshape(B, t, h, w, dt, dc).permute(0, 1, 4, 2, 3, 5).reshape(B, t * dt, h, w, dc)   # :882-886
    m = center_frame_mask(mask, cube_marker, dt, cfg['num_frames'])
    loss = ((p - target) ** 2).mean(dim=-1)                                          # :899-900
    loss = (loss * m).sum() / (m.sum() + 1e-5)                                       # :901
    return p, loss


# ----------------------------------------------------------------------------
# random state in the reference's key layout
# ----------------------------------------------------------------------------
def random_maskfeat_state(cfg: dict, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Random parameters under the reference's names and shapes; scales chosen so every term matters
    (pool weights, mask token and positional tables are non-trivial)."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=0.02):
        return (torch.randn(*shape, generator=g, dtype=torch.float64) * std).to(dtype)

    C0 = cfg['patch_embed_dim']
    T, H, W = cfg['thw']
    kt, kh, kw = cfg['kernel']
    sd = {
        'patch_embed.patch_model.weight': rn(C0, 3, kt, kh, kw, std=(2.0 / (3 * kt * kh * kw)) ** 0.5),
        'patch_embed.patch_model.bias': rn(C0),
        'mask_token': rn(1, 1, C0, std=0.5),
        'mvit.cls_positional_encoding.cls_token': rn(1, 1, C0, std=0.5),
        'mvit.cls_positional_encoding.pos_embed_spatial': rn(1, H * W, C0, std=0.2),
        'mvit.cls_positional_encoding.pos_embed_temporal': rn(1, T, C0, std=0.2),
        'mvit.cls_positional_encoding.pos_embed_class': rn(1, 1, C0, std=0.2),
    }
    for i, b in enumerate(cfg['blocks']):
        p = f'mvit.blocks.{i}.'
        d, do, hd = b['dim'], b['dim_out'], b['dim'] // b['heads']
        for n in ('norm1', 'norm2'):
            sd[p + n + '.weight'] = 1 + rn(d, std=0.1)
            sd[p + n + '.bias'] = rn(d, std=0.1)
        for n in ('q', 'k', 'v', 'proj'):
            sd[p + f'attn.{n}.weight'] = rn(d, d, std=d ** -0.5)
            sd[p + f'attn.{n}.bias'] = rn(d, std=0.02)
        pools = []
        if len(b['stride_q']) > 0:
            pools.append(('q', b['kernel_q']))
        if len(b['stride_kv']) > 0:
            pools += [('k', b['kernel_kv']), ('v', b['kernel_kv'])]
        for n, kern in pools:
            sd[p + f'attn.pool_{n}.weight'] = rn(hd, 1, *kern, std=(1.0 / math.prod(kern)) ** 0.5)
            sd[p + f'attn.norm_{n}.weight'] = 1 + rn(hd, std=0.1)
            sd[p + f'attn.norm_{n}.bias'] = rn(hd, std=0.1)
        sd[p + 'mlp.fc1.weight'] = rn(b['hidden'], d, std=d ** -0.5)
        sd[p + 'mlp.fc1.bias'] = rn(b['hidden'])
        sd[p + 'mlp.fc2.weight'] = rn(do, b['hidden'], std=b['hidden'] ** -0.5)
        sd[p + 'mlp.fc2.bias'] = rn(do)
        if d != do:
            sd[p + 'proj.weight'] = rn(do, d, std=d ** -0.5)
            sd[p + 'proj.bias'] = rn(do)
    D = cfg['out_dim']
    sd['mvit.norm_embed.weight'] = 1 + rn(D, std=0.1)
    sd['mvit.norm_embed.bias'] = rn(D, std=0.1)
    sd['decoder_pred.weight'] = rn(cfg['feature_dim'], D, std=D ** -0.5)
    sd['decoder_pred.bias'] = rn(cfg['feature_dim'])
    return sd
