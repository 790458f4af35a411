"""Initialisers used by the module surface (same names as the reference's weight_init.py:65-105).

The in-place init helpers, plus the checkpoint loaders the reference constructors call when `pretrain_pth` is given
(`init_from_vit_pretrain_` weight_init.py:107-186, `init_from_mae_pretrain_` :189-298, `init_from_kinetics_pretrain_`
:301-314).  The loaders are host-only key remapping (the package keeps the reference's state-dict keys), written here
as pure functions on dicts (`remap_*`) so they can be tested without files; `TimeSformer/ViViT/MaskFeat(pretrain_pth=...)`
route through them exactly like the reference (video_transformer.py:154-165, :434-451, :866-870).
"""
from __future__ import annotations

import math
import re

import torch
import torch.nn as nn


@torch.no_grad()
def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    """Truncated normal via inverse-CDF sampling (timm-style; absolute cut-offs a, b)."""
    cdf = lambda v: 0.5 * (1.0 + math.erf(v / math.sqrt(2.0)))
    lo, hi = cdf((a - mean) / std), cdf((b - mean) / std)
    tensor.uniform_(2 * lo - 1, 2 * hi - 1)
    tensor.erfinv_()
    tensor.mul_(std * math.sqrt(2.0))
    tensor.add_(mean)
    tensor.clamp_(min=a, max=b)
    return tensor


@torch.no_grad()
def constant_init_(tensor, constant_value=0):
    nn.init.constant_(tensor, constant_value)


@torch.no_grad()
def kaiming_init_(tensor, a=0, mode='fan_out', nonlinearity='relu', distribution='normal'):
    if distribution == 'uniform':
        nn.init.kaiming_uniform_(tensor, a=a, mode=mode, nonlinearity=nonlinearity)
    elif distribution == 'normal':
        nn.init.kaiming_normal_(tensor, a=a, mode=mode, nonlinearity=nonlinearity)
    else:
        raise ValueError(distribution)


# --------------------------------------------------------------------------------------------------
# checkpoint loaders (pretrain_pth)
# --------------------------------------------------------------------------------------------------
_LAYER_IDX = re.compile(r'(?<=layers\.)\d+')


def _inflate_patch_filter(weight, tube_size, extend_strategy):
    """Conv2d filter [D,C,h,w] -> Conv3d filter [D,C,tube,h,w] (weight_init.py:130-139).

    Quirk reproduced, not fixed: for 'center_frame' the reference zeroes `new_weight` in place, but einops.repeat returned
    an expanded VIEW of `weight`, so the 2-D filter is zeroed with it and the "centre frame" it then copies in is all
    zeros — the inflated filter is identically zero (checked against the reference function, tests/test_checkpoint_loaders.py)."""
    w3 = weight.unsqueeze(2).repeat(1, 1, tube_size, 1, 1)
    if extend_strategy == 'temporal_avg':
        w3 = w3 / tube_size
    elif extend_strategy == 'center_frame':
        w3 = torch.zeros_like(w3)
    return w3


def _replicate_attention(state, attention_type, copy_strategy, num_time_transformer_layers):
    """Second pass of both image-checkpoint loaders (weight_init.py:160-181): the image model has one attention per layer;
    divided space-time gets its spatial attention (`attentions.1`) from it, the factorised encoder fills its first
    `num_time_transformer_layers` temporal layers from the spatial ones — copied or zeroed."""
    for key in list(state.keys()):
        new_key = None
        if attention_type == 'divided_space_time':
            if 'attentions.0' in key:
                new_key = key.replace('attentions.0', 'attentions.1')
        elif attention_type == 'fact_encoder':
            idx = _LAYER_IDX.findall(key)
            if len(idx) > 1 and int(idx[1]) < num_time_transformer_layers:
                new_key = key.replace('transformer_layers.0.layers', 'transformer_layers.1.layers')
        if new_key is None:
            continue
        if copy_strategy == 'repeat':
            state[new_key] = state[key].clone()
        elif copy_strategy == 'set_zero':
            state[new_key] = torch.zeros_like(state[key])
    return state


def remap_vit_checkpoint(state, conv_type, attention_type, copy_strategy, extend_strategy='temporal_avg', tube_size=2,
                         num_time_transformer_layers=4):
    """mmaction-style ViT image checkpoint -> this package's keys (reference init_from_vit_pretrain_, :107-186)."""
    out = {}
    for key, val in state.items():
        if conv_type == 'Conv3d' and 'patch_embed.projection.weight' in key:
            out[key] = _inflate_patch_filter(val, tube_size, extend_strategy)
            continue
        nk = key.replace('transformer_layers.layers', 'transformer_layers.0.layers') if attention_type == 'fact_encoder' else key
        if 'in_proj' in nk:
            nk = nk.replace('in_proj_', 'qkv.')
        elif 'out_proj' in nk:
            nk = nk.replace('out_proj', 'proj')
        if 'norms' in nk:
            nk = nk.replace('norms.0', 'attentions.0.norm').replace('norms.1', 'ffns.0.norm')
        out[nk] = val
    return _replicate_attention(out, attention_type, copy_strategy, num_time_transformer_layers)


def remap_mae_checkpoint(state, conv_type, attention_type, copy_strategy, extend_strategy='temporal_avg', tube_size=2,
                         num_time_transformer_layers=4):
    """VideoMAE/BEiT-style encoder checkpoint -> this package's keys (reference init_from_mae_pretrain_, :189-298)."""
    out = {}
    blocks = 'transformer_layers.0.layers' if attention_type == 'fact_encoder' else 'transformer_layers.layers'
    for key, val in state.items():
        if 'decoder' in key:
            continue
        if 'encoder.patch_embed.proj' in key:
            nk = key.replace('encoder.patch_embed.proj', 'patch_embed.projection')
            if conv_type == 'Conv3d' and 'weight' in key:
                val = _inflate_patch_filter(val, tube_size, extend_strategy)
            out[nk] = val
            continue
        nk = key.replace('encoder.blocks', blocks)
        if 'norm' in nk:
            nk = nk.replace('norm1', 'attentions.0.norm').replace('norm2', 'ffns.0.norm')
        elif 'attn' in nk:
            if 'q_bias' in nk:                          # separate q / v biases, k has none: qkv bias = [q, 0, v]
                v_bias = state[key.replace('q_bias', 'v_bias')]
                out[nk.replace('attn.q_bias', 'attentions.0.attn.qkv.bias')] = torch.cat((val, torch.zeros_like(val), v_bias))
                continue
            if 'v_bias' in nk:
                continue
        elif 'mlp' in nk:
            nk = nk.replace('mlp.fc1', 'ffns.0.layers.0.0').replace('mlp.fc2', 'ffns.0.layers.1')
        if 'encoder.norm' in key:
            nk = key.replace('encoder.norm', 'norm')
        out[nk] = val
    return _replicate_attention(out, attention_type, copy_strategy, num_time_transformer_layers)


def remap_kinetics_checkpoint(state):
    """Lightning checkpoint of this very trainer (`model.*` / `cls_head.*` prefixes) -> module keys
    (reference replace_state_dict, weight_init.py:17-29)."""
    out = {}
    for key, val in state.items():
        if key.startswith('model'):
            nk = key[6:]
            if 'in_proj' in nk:
                nk = nk.replace('in_proj_', 'qkv.')
            elif 'out_proj' in nk:
                nk = nk.replace('out_proj', 'proj')
        else:
            nk = key[9:]
        out[nk] = val
    return out


def _read_checkpoint(pretrained, inner_key):
    state = torch.load(pretrained, map_location='cpu') if isinstance(pretrained, str) else dict(pretrained)
    return dict(state[inner_key]) if inner_key in state else state


@torch.no_grad()
def init_from_vit_pretrain_(module, pretrained, conv_type, attention_type, copy_strategy, extend_strategy='temporal_avg',
                            tube_size=2, num_time_transformer_layers=4):
    state = remap_vit_checkpoint(_read_checkpoint(pretrained, 'state_dict'), conv_type, attention_type, copy_strategy,
                                 extend_strategy, tube_size, num_time_transformer_layers)
    return module.load_state_dict(state, strict=False)


@torch.no_grad()
def init_from_mae_pretrain_(module, pretrained, conv_type, attention_type, copy_strategy, extend_strategy='temporal_avg',
                            tube_size=2, num_time_transformer_layers=4):
    state = remap_mae_checkpoint(_read_checkpoint(pretrained, 'model'), conv_type, attention_type, copy_strategy,
                                 extend_strategy, tube_size, num_time_transformer_layers)
    return module.load_state_dict(state, strict=False)


@torch.no_grad()
def init_from_kinetics_pretrain_(module, pretrain_pth):
    return module.load_state_dict(remap_kinetics_checkpoint(_read_checkpoint(pretrain_pth, 'state_dict')), strict=False)
