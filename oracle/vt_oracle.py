"""Plain-torch CPU restatement of the reference's TimeSformer / ViViT forward.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the
reference file:line it restates (paths relative to /root/reference).  The
restatement is functional: it consumes a reference-format ``state_dict`` and
uses only torch primitives (matmul, softmax, erf ...), no nn.Module and no
einops, so it is an independent statement of the algorithm.  The backward pass
is torch autograd over these primitives.

Pinned against the real reference modules by ``oracle/make_golden.py`` (run in
the build container, where /root/reference exists); the resulting vectors live
in ``tests/golden``.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

Tensor = torch.Tensor


# ----------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------
def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    """nn.LayerNorm over the last dim (biased variance), transformer.py:215,:321,:418,:495."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x: Tensor) -> Tensor:
    """nn.GELU() default = exact erf form, transformer.py:483."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    y = x @ w.t()
    return y if b is None else y + b


def drop_path(x: Tensor, p: float, training: bool) -> Tensor:
    """transformer.py:34-42 — mask per dim-0 row from the CPU default generator."""
    if p == 0.0 or not training:
        return x
    keep = 1.0 - p
    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
    r = (keep + torch.rand(shape).to(device=x.device, dtype=x.dtype)).floor()      # `.type_as(x)`: CPU draw, moved to x
    return x / keep * r


def attention_core(x: Tensor, sd: Dict[str, Tensor], pre: str, heads: int):
    """Attention.forward, transformer.py:165-177.  Returns (out, probs)."""
    Bp, N, C = x.shape
    hd = C // heads
    qkv = linear(x, sd[pre + 'qkv.weight'], sd[pre + 'qkv.bias'])       # :167
    qkv = qkv.reshape(Bp, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q @ k.transpose(-2, -1)) * (hd ** -0.5)                       # :170
    attn = torch.softmax(attn, dim=-1)                                    # :171
    out = (attn @ v).transpose(1, 2).reshape(Bp, N, C)                    # :174
    out = linear(out, sd[pre + 'proj.weight'], sd[pre + 'proj.bias'])     # :175
    return out, attn


# ----------------------------------------------------------------------------
# blocks
# ----------------------------------------------------------------------------
def divided_temporal(x, sd, pre, T, heads, dp, training, return_attention=False):
    """DividedTemporalAttentionWithPreNorm.forward with use_cls_token=False,
    transformer.py:234-282 (cls bypasses the block, :281)."""
    B, N1, D = x.shape
    cls = x[:, :1, :]
    q = x[:, 1:, :]
    res = q
    P = (N1 - 1) // T
    q = q.reshape(B * P, T, D)                                            # :250  'b (p t) d -> (b p) t d'
    q = layer_norm(q, sd[pre + 'norm.weight'], sd[pre + 'norm.bias'], 1e-5)
    out, attn = attention_core(q, sd, pre + 'attn.', heads)
    if return_attention:
        return attn
    out = drop_path(out, dp, training)                                    # :265
    out = linear(out, sd[pre + 'temporal_fc.weight'], sd[pre + 'temporal_fc.bias'])  # :267
    out = out.reshape(B, P * T, D)                                        # :279
    return torch.cat((cls, res + out), dim=1)                             # :280-281


def divided_spatial(x, sd, pre, T, heads, dp, training, return_attention=False):
    """DividedSpatialAttentionWithPreNorm.forward with use_cls_token=True,
    transformer.py:336-382."""
    B, N1, D = x.shape
    res = x
    cls = x[:, :1, :]
    q = x[:, 1:, :]
    P = (N1 - 1) // T
    q = q.reshape(B, P, T, D).permute(0, 2, 1, 3).reshape(B * T, P, D)    # :352  'b (p t) d -> (b t) p d'
    cls_rep = cls.expand(B, T, D).reshape(B * T, 1, D)                    # :354-355
    q = torch.cat((cls_rep, q), dim=1)                                    # :356
    q = layer_norm(q, sd[pre + 'norm.weight'], sd[pre + 'norm.bias'], 1e-5)
    out, attn = attention_core(q, sd, pre + 'attn.', heads)
    if return_attention:
        return attn
    out = drop_path(out, dp, training)                                    # :367
    cls_out = out[:, 0, :].reshape(B, T, D).mean(dim=1, keepdim=True)     # :371-373
    out = out[:, 1:, :].reshape(B, T, P, D).permute(0, 2, 1, 3).reshape(B, P * T, D)  # :375
    return res + torch.cat((cls_out, out), dim=1)                         # :376-377


def mha_prenorm(x, sd, pre, heads, dp, training, return_attention=False):
    """MultiheadAttentionWithPreNorm.forward, transformer.py:428-456."""
    q = layer_norm(x, sd[pre + 'norm.weight'], sd[pre + 'norm.bias'], 1e-5)
    out, attn = attention_core(q, sd, pre + 'attn.', heads)
    if return_attention:
        return attn
    return x + drop_path(out, dp, training)


def ffn_prenorm(x, sd, pre, dp, training):
    """FFNWithPreNorm.forward, transformer.py:516-523 (layers :496-507)."""
    h = layer_norm(x, sd[pre + 'norm.weight'], sd[pre + 'norm.bias'], 1e-5)
    h = linear(h, sd[pre + 'layers.0.0.weight'], sd[pre + 'layers.0.0.bias'])
    h = gelu_erf(h)
    h = linear(h, sd[pre + 'layers.1.weight'], sd[pre + 'layers.1.bias'])
    return x + drop_path(h, dp, training)


def container(x, sd, pre, num_layers, order, T, heads, training,
              drop_path_rate=0.1, return_attention=False):
    """TransformerContainer / BasicTransformerBlock, transformer.py:526-636.
    dpr = linspace(0, rate, L) (:543); every sub-block of layer i shares dpr[i]."""
    dpr = np.linspace(0, drop_path_rate, num_layers)
    n_attn = sum(1 for o in order if o != 'ffn')
    for i in range(num_layers):
        lp = f'{pre}layers.{i}.'
        last = return_attention and i >= num_layers - 1                   # :560
        a = 0
        for op in order:
            if op == 'ffn':
                continue
            want = last and a >= n_attn - 1                               # :628
            ap = f'{lp}attentions.{a}.'
            if op == 'time_attn':
                x = divided_temporal(x, sd, ap, T, heads, float(dpr[i]), training, want)
            elif op == 'space_attn':
                x = divided_spatial(x, sd, ap, T, heads, float(dpr[i]), training, want)
            elif op == 'self_attn':
                x = mha_prenorm(x, sd, ap, heads, float(dpr[i]), training, want)
            else:
                raise TypeError(op)
            if want:
                return x
            a += 1
        x = ffn_prenorm(x, sd, f'{lp}ffns.0.', float(dpr[i]), training)
    return x


# ----------------------------------------------------------------------------
# patch / tubelet embedding
# ----------------------------------------------------------------------------
def patch_embed(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """PatchEmbed.forward, transformer.py:138-151.  Non-overlapping conv == unfold + matmul.
    Conv2d weight (D,3,ph,pw) on '(b t) c h w'; Conv3d weight (D,3,tube,ph,pw) on 'b c t h w'.
    Returns ((b t'), (h w), D)."""
    B, T, C, H, W = x.shape
    if w.ndim == 4:
        tube, ph, pw = 1, w.shape[2], w.shape[3]
    else:
        tube, ph, pw = w.shape[2], w.shape[3], w.shape[4]
    D = w.shape[0]
    Tp, Hp, Wp = T // tube, H // ph, W // pw
    # (B, Tp, tube, C, Hp, ph, Wp, pw) -> (B, Tp, Hp, Wp, C, tube, ph, pw)
    xx = x.reshape(B, Tp, tube, C, Hp, ph, Wp, pw).permute(0, 1, 4, 6, 3, 2, 5, 7)
    xx = xx.reshape(B * Tp, Hp * Wp, C * tube * ph * pw)
    return xx @ w.reshape(D, -1).t() + b


# ----------------------------------------------------------------------------
# models
# ----------------------------------------------------------------------------
def timesformer_tokens(sd, x, cfg):
    """TimeSformer.prepare_tokens (divided_space_time), video_transformer.py:193-240."""
    B = x.shape[0]
    tok = patch_embed(x, sd['patch_embed.projection.weight'], sd['patch_embed.projection.bias'])
    BT, P, D = tok.shape
    T = BT // B                                                           # frames, or tubelets for a Conv3d embedding
    cls = sd['cls_token'].expand(BT, 1, D)
    tok = torch.cat((cls, tok), dim=1) + sd['pos_embed']                  # :207-209
    cls_tokens = tok[:B, 0, :].unsqueeze(1)                               # :216
    tok = tok[:, 1:, :].reshape(B, T, P, D).permute(0, 2, 1, 3).reshape(B * P, T, D)  # :231
    tok = tok + sd['time_embed']                                          # :233
    tok = tok.reshape(B, P * T, D)                                        # :236
    return torch.cat((cls_tokens, tok), dim=1)                            # :237


def timesformer_forward(sd, x, cfg, training=False, return_tokens=False):
    """TimeSformer.forward, video_transformer.py:242-256 (attention_type=divided_space_time)."""
    tok = timesformer_tokens(sd, x, cfg)
    tok = container(tok, sd, 'transformer_layers.', cfg['num_transformer_layers'],
                    ['time_attn', 'space_attn', 'ffn'], cfg['num_frames'], cfg['num_heads'],
                    training)
    tok = layer_norm(tok, sd['norm.weight'], sd['norm.bias'], 1e-6)       # :251, eps :119
    if return_tokens:
        return tok
    return tok[:, 0]                                                      # :254


def timesformer_space_only_forward(sd, x, cfg, training=False):
    """TimeSformer.forward with attention_type='space_only' (video_transformer.py:193-212, :242-256): per-frame
    tokens (no time embedding), joint attention per frame, mean over frames, final norm, cls."""
    B, T = x.shape[0], x.shape[1]
    tok = patch_embed(x, sd['patch_embed.projection.weight'], sd['patch_embed.projection.bias'])
    BT, P, D = tok.shape
    tok = torch.cat((sd['cls_token'].expand(BT, 1, D), tok), dim=1) + sd['pos_embed']          # :207-209
    tok = container(tok, sd, 'transformer_layers.', cfg['num_transformer_layers'], ['self_attn', 'ffn'],
                    cfg['num_frames'], cfg['num_heads'], training)
    tok = tok.reshape(B, T, P + 1, D).mean(dim=1)                                               # :248-249
    tok = layer_norm(tok, sd['norm.weight'], sd['norm.bias'], 1e-6)
    return tok[:, 0]


def timesformer_joint_forward(sd, x, cfg, training=False):
    """TimeSformer.forward with attention_type='joint_space_time' (video_transformer.py:104-116, :193-256): the tokens
    of the divided variant (spatial + temporal embedding, one cls), then ['self_attn', 'ffn'] layers that attend over
    all 1 + P*T tokens of a clip at once."""
    tok = timesformer_tokens(sd, x, cfg)
    tok = container(tok, sd, 'transformer_layers.', cfg['num_transformer_layers'], ['self_attn', 'ffn'],
                    cfg['num_frames'], cfg['num_heads'], training)
    tok = layer_norm(tok, sd['norm.weight'], sd['norm.bias'], 1e-6)
    return tok[:, 0]


def timesformer_last_selfattention(sd, x, cfg):
    """TimeSformer.get_last_selfattention, video_transformer.py:258-261."""
    tok = timesformer_tokens(sd, x, cfg)
    return container(tok, sd, 'transformer_layers.', cfg['num_transformer_layers'],
                     ['time_attn', 'space_attn', 'ffn'], cfg['num_frames'], cfg['num_heads'],
                     False, return_attention=True)


def vivit_forward(sd, x, cfg, training=False):
    """ViViT.forward (fact_encoder), video_transformer.py:455-532, incl. the
    `x[:b,0,:]` cls-gather quirk at :515."""
    B = x.shape[0]
    tok = patch_embed(x, sd['patch_embed.projection.weight'], sd['patch_embed.projection.bias'])
    BT, P, D = tok.shape
    Tp = BT // B
    cls = sd['cls_token'].expand(BT, 1, D)
    tok = torch.cat((cls, tok), dim=1) + sd['pos_embed']                  # :469-471
    tok = container(tok, sd, 'transformer_layers.0.', cfg['num_transformer_layers'],
                    ['self_attn', 'ffn'], Tp, cfg['num_heads'], training)  # :512
    cls_tokens = tok[:B, 0, :].unsqueeze(1)                               # :515 (quirk: rows 0..B-1 of (b t))
    t = tok[:, 1:, :].reshape(B, Tp, P, D).mean(dim=2)                    # :516-517
    t = torch.cat((cls_tokens, t), dim=1) + sd['time_embed']              # :518-520
    t = container(t, sd, 'transformer_layers.1.', 4, ['self_attn', 'ffn'], Tp,
                  cfg['num_heads'], training)                             # :525, 4 layers :377
    t = layer_norm(t, sd['norm.weight'], sd['norm.bias'], 1e-6)           # :527
    return t[:, 0]                                                        # :530


def vivit_variant_forward(sd, x, cfg, attention_type, training=False):
    """ViViT.forward with attention_type 'joint_space_time' (model 1) or 'divided_space_time' (model 3),
    video_transformer.py:349-373, :455-510, :527-532.  With `use_cls_token_temporal` False for both (:405-411) the token
    assembly (:461-499) is TimeSformer's, on tubelets: cls + pos_embed(1+P), time_embed(T') per patch, tokens 'b (p t) d'."""
    tok = timesformer_tokens(sd, x, cfg)
    Tp = (tok.shape[1] - 1) // ((cfg['img_size'] // cfg['patch_size']) ** 2)
    order = ['self_attn', 'ffn'] if attention_type == 'joint_space_time' else ['time_attn', 'space_attn', 'ffn']
    tok = container(tok, sd, 'transformer_layers.', cfg['num_transformer_layers'], order, Tp, cfg['num_heads'], training)
    tok = layer_norm(tok, sd['norm.weight'], sd['norm.bias'], 1e-6)
    return tok[:, 0]


def classification_head(sd, x, pre='cls_head.'):
    """ClassificationHead.forward, transformer.py:78-80."""
    return linear(x, sd[pre + 'weight'], sd[pre + 'bias'])


# ----------------------------------------------------------------------------
# random reference-format state dicts (for CPU baselines on boxes without /root/reference)
# ----------------------------------------------------------------------------
def random_timesformer_state(cfg, seed=0, dtype=torch.float32, randomize_temporal_fc=True):
    """Shapes/keys of SURVEY.md A.4.  Distributions approximate the reference init
    (A.3); `temporal_fc` gets N(0,0.02) instead of the reference's zeros so the
    temporal branch is exercised."""
    g = torch.Generator().manual_seed(seed)
    D, H, L, T = cfg['embed_dims'], cfg['num_heads'], cfg['num_transformer_layers'], cfg['num_frames']
    ps = cfg['patch_size']
    P = (cfg['img_size'] // ps) ** 2
    C = cfg.get('in_channels', 3)

    def rn(*s, std=0.02):
        return (torch.randn(*s, generator=g, dtype=torch.float64) * std).to(dtype)

    def lin(o, i):
        bound = 1.0 / math.sqrt(i)
        w = (torch.rand(o, i, generator=g, dtype=torch.float64) * 2 - 1) * bound
        b = (torch.rand(o, generator=g, dtype=torch.float64) * 2 - 1) * bound
        return w.to(dtype), b.to(dtype)

    sd = {}
    sd['cls_token'] = rn(1, 1, D)
    sd['pos_embed'] = rn(1, P + 1, D)
    sd['time_embed'] = rn(1, T, D)
    fan_in = C * ps * ps
    sd['patch_embed.projection.weight'] = rn(D, C, ps, ps, std=math.sqrt(2.0 / fan_in))
    sd['patch_embed.projection.bias'] = rn(D, std=0.02)
    sd['norm.weight'] = (1 + rn(D, std=0.1))
    sd['norm.bias'] = rn(D, std=0.1)
    for i in range(L):
        for a in (0, 1):
            p = f'transformer_layers.layers.{i}.attentions.{a}.'
            sd[p + 'norm.weight'] = 1 + rn(D, std=0.1)
            sd[p + 'norm.bias'] = rn(D, std=0.1)
            sd[p + 'attn.qkv.weight'], sd[p + 'attn.qkv.bias'] = lin(3 * D, D)
            sd[p + 'attn.proj.weight'], sd[p + 'attn.proj.bias'] = lin(D, D)
            if a == 0:
                if randomize_temporal_fc:
                    sd[p + 'temporal_fc.weight'] = rn(D, D)
                    sd[p + 'temporal_fc.bias'] = rn(D)
                else:
                    sd[p + 'temporal_fc.weight'] = torch.zeros(D, D, dtype=dtype)
                    sd[p + 'temporal_fc.bias'] = torch.zeros(D, dtype=dtype)
        p = f'transformer_layers.layers.{i}.ffns.0.'
        sd[p + 'norm.weight'] = 1 + rn(D, std=0.1)
        sd[p + 'norm.bias'] = rn(D, std=0.1)
        sd[p + 'layers.0.0.weight'], sd[p + 'layers.0.0.bias'] = lin(4 * D, D)
        sd[p + 'layers.1.weight'], sd[p + 'layers.1.bias'] = lin(D, 4 * D)
    return sd


TIMESFORMER_B = dict(num_frames=8, img_size=224, patch_size=16, embed_dims=768, num_heads=12,
                     num_transformer_layers=12, attention_type='divided_space_time')
