"""nn.Module shells around ``oracle/mvit_oracle.py`` under pytorchvideo's class and parameter names.

TEST INFRASTRUCTURE.  pytorchvideo is not installed in this image and not vendored by the reference, yet the
reference's ``video_transformer.py:15-17`` imports five names from it.  ``oracle/make_golden.py`` binds those names to
the classes below so that the REAL ``MaskFeat`` (and its in-tree factory ``create_multiscale_vision_transformers``,
video_transformer.py:621-800) can be constructed and run: everything the reference itself implements (block
configuration, conv patch embed, mask-token mixing, decoder, loss) is then pinned by the reference's own code, and only
the block arithmetic comes from the restatement (parity unpinned for that part, see mvit_oracle.py).

Constructor signatures follow the keyword set the reference passes (video_transformer.py:764-785, :681-689, :793-800).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import mvit_oracle as mo


def round_width(width, multiplier, min_width=8, divisor=8, ceil=False):
    assert not ceil
    return mo.round_width(width, multiplier, min_width=min_width, divisor=divisor)


def set_attributes(self, params=None):
    """pytorchvideo.layers.utils.set_attributes: copy constructor locals onto the module."""
    if params:
        for k, v in params.items():
            if k != 'self':
                setattr(self, k, v)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features, out_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, out_features)


class MultiScaleAttention(nn.Module):
    def __init__(self, dim, num_heads, qkv_bias, kernel_q, kernel_kv, stride_q, stride_kv, norm_layer):
        super().__init__()
        hd = dim // num_heads
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.k = nn.Linear(dim, dim, bias=qkv_bias)
        self.v = nn.Linear(dim, dim, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

        def pool(kernel, stride):
            if len(stride) == 0 or (math.prod(kernel) == 1 and math.prod(stride) == 1):
                return None, None
            conv = nn.Conv3d(hd, hd, tuple(kernel), stride=tuple(stride), padding=tuple(k // 2 for k in kernel),
                             groups=hd, bias=False)
            return conv, norm_layer(hd)

        self.pool_q, self.norm_q = pool(kernel_q, stride_q)
        self.pool_k, self.norm_k = pool(kernel_kv, stride_kv)
        self.pool_v, self.norm_v = pool(kernel_kv, stride_kv)


class MultiScaleBlock(nn.Module):
    def __init__(self, dim, dim_out, num_heads, mlp_ratio=4.0, qkv_bias=False, dropout_rate=0.0, droppath_rate=0.0,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, kernel_q=(1, 1, 1), kernel_kv=(1, 1, 1),
                 stride_q=(1, 1, 1), stride_kv=(1, 1, 1), pool_mode='conv', has_cls_embed=True, pool_first=False):
        super().__init__()
        assert pool_mode == 'conv' and has_cls_embed and not pool_first and dropout_rate == 0.0 and droppath_rate == 0.0
        self.blk = dict(dim=dim, dim_out=dim_out, heads=num_heads, kernel_q=list(kernel_q), stride_q=list(stride_q),
                        kernel_kv=list(kernel_kv), stride_kv=list(stride_kv), hidden=int(dim * mlp_ratio))
        self.norm1 = norm_layer(dim)
        self.attn = MultiScaleAttention(dim, num_heads, qkv_bias, kernel_q, kernel_kv, stride_q, stride_kv,
                                        nn.LayerNorm)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio), dim_out)
        if dim != dim_out:
            self.proj = nn.Linear(dim, dim_out)
        self.eps_block = self.norm1.eps
        self.eps_pool = 1e-5

    def forward(self, x, thw):
        sd = dict(self.named_parameters())
        y, new_thw = mo.multiscale_block(sd, '', x, thw, self.blk, self.eps_block, self.eps_pool)
        return y, list(new_thw)


class SpatioTemporalClsPositionalEncoding(nn.Module):
    def __init__(self, embed_dim, patch_embed_shape, sep_pos_embed=False, has_cls=True):
        super().__init__()
        assert sep_pos_embed and has_cls
        self.patch_embed_shape = list(patch_embed_shape)
        T, H, W = patch_embed_shape
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed_spatial = nn.Parameter(torch.zeros(1, H * W, embed_dim))
        self.pos_embed_temporal = nn.Parameter(torch.zeros(1, T, embed_dim))
        self.pos_embed_class = nn.Parameter(torch.zeros(1, 1, embed_dim))

    def forward(self, x):
        return mo.cls_positional_encoding(dict(self.named_parameters()), '', x, self.patch_embed_shape)


class MultiscaleVisionTransformers(nn.Module):
    def __init__(self, *, patch_embed, cls_positional_encoding, pos_drop, norm_patch_embed, blocks, norm_embed, head):
        super().__init__()
        assert patch_embed is None and pos_drop is None and norm_patch_embed is None and head is None
        self.cls_positional_encoding = cls_positional_encoding
        self.blocks = blocks
        self.norm_embed = norm_embed

    def forward(self, x):
        x = self.cls_positional_encoding(x)
        thw = self.cls_positional_encoding.patch_embed_shape
        for blk in self.blocks:
            x, thw = blk(x, thw)
        return self.norm_embed(x)


def install_stub_modules():
    """Register ``pytorchvideo.layers[.utils]`` / ``pytorchvideo.models.vision_transformers`` in sys.modules."""
    import sys
    import types
    pv = types.ModuleType('pytorchvideo')
    layers = types.ModuleType('pytorchvideo.layers')
    utils = types.ModuleType('pytorchvideo.layers.utils')
    models = types.ModuleType('pytorchvideo.models')
    vts = types.ModuleType('pytorchvideo.models.vision_transformers')
    layers.MultiScaleBlock = MultiScaleBlock
    layers.SpatioTemporalClsPositionalEncoding = SpatioTemporalClsPositionalEncoding
    layers.utils = utils
    utils.round_width = round_width
    utils.set_attributes = set_attributes
    vts.MultiscaleVisionTransformers = MultiscaleVisionTransformers
    models.vision_transformers = vts
    pv.layers, pv.models = layers, models
    for name, mod in (('pytorchvideo', pv), ('pytorchvideo.layers', layers), ('pytorchvideo.layers.utils', utils),
                      ('pytorchvideo.models', models), ('pytorchvideo.models.vision_transformers', vts)):
        sys.modules[name] = mod
