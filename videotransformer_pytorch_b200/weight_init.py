"""Initialisers used by the module surface (same names as the reference's weight_init.py:65-105).

Only the in-place init helpers live here.  The reference's checkpoint key-remapping functions
(`init_from_vit_pretrain_` etc., weight_init.py:107-314) are control-plane code outside the hot path
(SURVEY.md §2.1): they operate purely on state-dict keys, which this package keeps identical, so the
reference's own functions can be applied to these modules unchanged.
"""
from __future__ import annotations

import math

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
