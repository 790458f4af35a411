"""nn.Module surface of the transformer building blocks — same class names, constructor arguments,
parameter names/shapes and forward signatures as the reference's transformer.py, so state dicts
round-trip with strict=True and model_trainer.py / optimizer.py / weight_init.py drive them unchanged.
Every forward is routed to the sm_100a kernels through ops.py; there is no eager/CPU fallback.

Reference classes mirrored (file:line in the reference repo):
  DropPath :25, ClassificationHead :45, PatchEmbed :83, Attention :153,
  DividedTemporalAttentionWithPreNorm :179, DividedSpatialAttentionWithPreNorm :285,
  MultiheadAttentionWithPreNorm :385, FFNWithPreNorm :459, TransformerContainer :526,
  BasicTransformerBlock :568, get_sine_cosine_pos_emb :12.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair

from . import _lib, ops
from .weight_init import constant_init_, kaiming_init_, trunc_normal_


def get_sine_cosine_pos_emb(n_position, d_hid):
    """Sinusoid table (1, n_position, d_hid): even dims sin, odd dims cos of pos / 10000^(2*(j//2)/d)."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid, dtype=np.float64)[None, :]
    angle = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table = np.where((np.arange(d_hid) % 2 == 0)[None, :], np.sin(angle), np.cos(angle))
    return torch.FloatTensor(table).unsqueeze(0)


class ShadowWeights:
    """bf16 shadows of fp32 parameters, refreshed when the parameter is modified (optimizer step,
    load_state_dict).  Owned by the module (SURVEY.md §8b 'Ownership')."""

    def __init__(self):
        self._cache = {}

    def get(self, name: str, param: torch.Tensor) -> torch.Tensor:
        key = (param._version, param.data_ptr(), param.device)
        hit = self._cache.get(name)
        capturing = param.is_cuda and torch.cuda.is_current_stream_capturing()
        if hit is not None and hit[0] == key and not capturing:
            return hit[1]
        with torch.no_grad():
            w = param.detach()
            if w.dtype != torch.float32:
                w = w.float()
            # while a CUDA graph is being captured the cast is always issued, so every replay re-derives the
            # shadow from the current fp32 values (the optimizer runs between replays)
            sh = _lib.K.cast_bf16(w.reshape(w.shape[0], -1).contiguous())
        self._cache[name] = (key, sh)
        return sh


    def get_cat(self, name: str, params) -> torch.Tensor:
        """bf16 shadow of several [n_i, k] weights stacked along dim 0 (fused q/k/v projection)."""
        key = tuple((p._version, p.data_ptr(), p.device) for p in params)
        hit = self._cache.get(name)
        capturing = params[0].is_cuda and torch.cuda.is_current_stream_capturing()
        if hit is not None and hit[0] == key and not capturing:
            return hit[1]
        with torch.no_grad():
            w = torch.cat([_f32(p.detach()).reshape(p.shape[0], -1) for p in params], dim=0)
            sh = _lib.K.cast_bf16(w.contiguous())
        self._cache[name] = (key, sh)
        return sh

    def get_padded(self, name: str, param: torch.Tensor, kpad: int) -> torch.Tensor:
        """bf16 shadow of a [n, ...] weight flattened to [n, k] and zero-padded to kpad columns (TMA needs 16-byte
        row pitches; the 3x7x7x3 patch-embed filter has k = 441)."""
        key = (param._version, param.data_ptr(), param.device, kpad)
        hit = self._cache.get(name)
        capturing = param.is_cuda and torch.cuda.is_current_stream_capturing()
        if hit is not None and hit[0] == key and not capturing:
            return hit[1]
        with torch.no_grad():
            w = _f32(param.detach()).reshape(param.shape[0], -1)
            w = torch.nn.functional.pad(w, (0, kpad - w.shape[1]))
            sh = _lib.K.cast_bf16(w.contiguous())
        self._cache[name] = (key, sh)
        return sh


def _f32(p):
    return p if p.dtype == torch.float32 else p.float()


class DropPath(nn.Module):
    """Stochastic depth; mask per dim-0 row from the CPU generator exactly like the reference
    (transformer.py:34-42).  Inside the fused blocks the same draw feeds the GEMM epilogue's row scale;
    this module's own forward is used when DropPath is applied to a stand-alone tensor."""

    def __init__(self, dropout_p=None):
        super().__init__()
        self.dropout_p = dropout_p

    def row_scale(self, n0, repeat, device):
        return ops.drop_path_scale(float(self.dropout_p or 0.), self.training, n0, repeat, device)

    def forward(self, x):
        s = self.row_scale(x.shape[0], 1, x.device)
        if s is None:
            return x
        return x * s.view((-1,) + (1,) * (x.ndim - 1)).to(x.dtype)


def _make_layer_drop(layer_drop):
    # reference builds fresh dicts per block and pops them (transformer.py:221-223); do not mutate the argument
    if not layer_drop:
        return nn.Identity()
    cfg = dict(layer_drop)
    p = cfg.pop('dropout_p')
    cls = cfg.pop('type')
    return cls(p) if cls else nn.Identity()


def _dp_scale(layer_drop, n0, repeat, device):
    return layer_drop.row_scale(n0, repeat, device) if isinstance(layer_drop, DropPath) else None


class ClassificationHead(nn.Module):
    """Linear classifier on the cls feature (reference transformer.py:45-80).  `cls_head` holds the parameters under the
    reference's names; forward is the fp32 skinny-GEMV kernel (ops.LinearSmallFn), `loss` adds the fused softmax-CE kernel."""

    def __init__(self, num_classes, in_channels, init_std=0.02, eval_metrics='finetune', **kwargs):
        super().__init__()
        self.init_std = init_std
        self.eval_metrics = eval_metrics
        self.cls_head = nn.Linear(in_channels, num_classes)
        self.init_weights(self.cls_head)

    def init_weights(self, module):
        if getattr(module, 'weight', None) is not None:
            if self.eval_metrics == 'finetune':
                trunc_normal_(module.weight, std=self.init_std)
            else:
                module.weight.data.normal_(mean=0.0, std=0.01)
        if getattr(module, 'bias', None) is not None:
            constant_init_(module.bias, constant_value=0)

    def forward(self, x):
        return ops.LinearSmallFn.apply(x, _f32(self.cls_head.weight), None if self.cls_head.bias is None else _f32(self.cls_head.bias))

    def loss(self, x, target):
        """mean cross-entropy of the head's logits: int64 labels (nn.CrossEntropyLoss, model_trainer.py:91, :207-208) or
        soft targets from Mixup (SoftTargetCrossEntropy, :89)."""
        return ops.cross_entropy(self.forward(x), target)


class PatchEmbed(nn.Module):
    """Non-overlapping patch (Conv2d) / tubelet (Conv3d) projection == im2col + tcgen05 GEMM.
    Stand-alone forward returns ((b t'), (h w), D) like the reference; the models call
    ops.PatchTokensFn, which fuses the positional/temporal embedding and the token regroup."""

    def __init__(self, img_size, patch_size, tube_size=2, in_channels=3, embed_dims=768, conv_type='Conv2d'):
        super().__init__()
        self.img_size = _pair(img_size)
        self.patch_size = _pair(patch_size)
        self.num_patches = (self.img_size[1] // self.patch_size[1]) * (self.img_size[0] // self.patch_size[0])
        if conv_type == 'Conv2d':
            self.projection = nn.Conv2d(in_channels, embed_dims, kernel_size=patch_size, stride=patch_size)
        elif conv_type == 'Conv3d':
            self.projection = nn.Conv3d(in_channels, embed_dims, kernel_size=(tube_size, patch_size, patch_size),
                                        stride=(tube_size, patch_size, patch_size))
        else:
            raise TypeError(f'Unsupported conv layer type {conv_type}')
        self.init_weights(self.projection)
        self._shadow = ShadowWeights()

    def init_weights(self, module):
        if getattr(module, 'weight', None) is not None:
            kaiming_init_(module.weight, mode='fan_in', nonlinearity='relu')
        if getattr(module, 'bias', None) is not None:
            constant_init_(module.bias, constant_value=0)

    @property
    def tube(self):
        return self.projection.kernel_size[0] if isinstance(self.projection, nn.Conv3d) else 1

    def shadow(self):
        return self._shadow.get('w', self.projection.weight)

    def forward(self, x):
        B, T = x.shape[0], x.shape[1]
        D = self.projection.weight.shape[0]
        dev = x.device
        zeros = lambda *s: torch.zeros(*s, device=dev)
        tok = ops.PatchTokensFn.apply(x, _f32(self.projection.weight), _f32(self.projection.bias), zeros(1, 1, D),
                                      zeros(1, self.num_patches + 1, D), None, self.shadow(), 'frames', self.tube)
        return tok[:, 1:, :]


class Attention(nn.Module):
    """qkv Linear -> softmax(q k^T * scale) v -> proj Linear; returns (out, attn) like the reference."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = qk_scale or head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self._shadow = ShadowWeights()
        if attn_drop or proj_drop:
            raise NotImplementedError('attn_drop / proj_drop > 0 are not on the reference hot path (always 0.)')

    def shadows(self):
        return self._shadow.get('qkv', self.qkv.weight), self._shadow.get('proj', self.proj.weight)

    def qkv_bias_or_zeros(self):
        if self.qkv.bias is not None:
            return _f32(self.qkv.bias)
        return torch.zeros(self.qkv.weight.shape[0], device=self.qkv.weight.device)

    def forward(self, x, need_weights=True):
        qh, ph = self.shadows()
        out, attn = ops.AttentionCoreFn.apply(x, _f32(self.qkv.weight), self.qkv_bias_or_zeros(), _f32(self.proj.weight),
                                              _f32(self.proj.bias), qh, ph, self.num_heads, need_weights)
        return out, (attn if need_weights else None)


class _DividedBase(nn.Module):
    def __init__(self, embed_dims, num_heads, num_frames, use_cls_token, attn_drop=0., proj_drop=0.,
                 layer_drop=None, norm_layer=nn.LayerNorm, **kwargs):
        super().__init__()
        if layer_drop is None:
            layer_drop = dict(type=DropPath, dropout_p=0.1)
        self.embed_dims = embed_dims
        self.num_heads = num_heads
        self.num_frames = num_frames
        self.use_cls_token = use_cls_token
        self.norm = norm_layer(embed_dims)
        self.attn = Attention(embed_dims, num_heads, qkv_bias=True, attn_drop=attn_drop)
        self.proj_drop = nn.Dropout(proj_drop)
        self.layer_drop = _make_layer_drop(layer_drop)
        if proj_drop:
            raise NotImplementedError('proj_drop > 0 is not on the reference hot path')


class DividedTemporalAttentionWithPreNorm(_DividedBase):
    """Temporal pass of divided space-time attention (attention over the T frames of each patch).
    Hot-path configuration: use_cls_token=False (cls bypasses the block, `temporal_fc` after DropPath)."""

    def __init__(self, embed_dims, num_heads, num_frames, use_cls_token, attn_drop=0., proj_drop=0.,
                 layer_drop=None, norm_layer=nn.LayerNorm, **kwargs):
        super().__init__(embed_dims, num_heads, num_frames, use_cls_token, attn_drop, proj_drop, layer_drop,
                         norm_layer, **kwargs)
        if not use_cls_token:
            self.temporal_fc = nn.Linear(self.embed_dims, self.embed_dims)
            self.init_weights(self.temporal_fc)

    def init_weights(self, module):
        # zero-init: the temporal branch starts as identity (reference transformer.py:228-232)
        if getattr(module, 'weight', None) is not None:
            constant_init_(module.weight, constant_value=0)
        if getattr(module, 'bias', None) is not None:
            constant_init_(module.bias, constant_value=0)

    def forward(self, query, key=None, value=None, residual=None, return_attention=False, **kwargs):
        assert residual is None, 'Always adding the shortcut in the forward function'
        if self.use_cls_token:
            raise NotImplementedError('temporal attention with use_cls_token=True is not used by any reference model')
        x = _f32(query).contiguous()
        B, S, D = x.shape
        T = self.num_frames
        P = (S - 1) // T
        if return_attention:
            maps = ops.token_maps(B, T, P, str(x.device))
            xn = ops.RowsNormFn.apply(x, _f32(self.norm.weight), _f32(self.norm.bias), self.norm.eps, maps['temporal'])
            return self.attn(xn.view(B * P, T, D))[1]
        dp = _dp_scale(self.layer_drop, B * P, T, x.device)
        qh, ph = self.attn.shadows()
        fh = self.attn._shadow.get('temporal_fc', self.temporal_fc.weight)
        return ops.TemporalAttnFn.apply(
            x, _f32(self.norm.weight), _f32(self.norm.bias), _f32(self.attn.qkv.weight), _f32(self.attn.qkv.bias),
            _f32(self.attn.proj.weight), _f32(self.attn.proj.bias), _f32(self.temporal_fc.weight),
            _f32(self.temporal_fc.bias), qh, ph, fh, dp, T, self.num_heads, self.norm.eps)


class DividedSpatialAttentionWithPreNorm(_DividedBase):
    """Spatial pass (attention over cls + the P patches of each frame, cls averaged over frames).
    Hot-path configuration: use_cls_token=True."""

    def __init__(self, embed_dims, num_heads, num_frames, use_cls_token, attn_drop=0., proj_drop=0.,
                 layer_drop=None, norm_layer=nn.LayerNorm, **kwargs):
        super().__init__(embed_dims, num_heads, num_frames, use_cls_token, attn_drop, proj_drop, layer_drop,
                         norm_layer, **kwargs)
        self.init_weights()

    def init_weights(self):
        pass

    def forward(self, query, key=None, value=None, residual=None, return_attention=False, **kwargs):
        assert residual is None, 'Always adding the shortcut in the forward function'
        if not self.use_cls_token:
            raise NotImplementedError('spatial attention with use_cls_token=False is not used by any reference model')
        x = _f32(query).contiguous()
        B, S, D = x.shape
        T = self.num_frames
        P = (S - 1) // T
        if return_attention:
            maps = ops.token_maps(B, T, P, str(x.device))
            xn = ops.RowsNormFn.apply(x, _f32(self.norm.weight), _f32(self.norm.bias), self.norm.eps, maps['sp_in'])
            return self.attn(xn.view(B * T, P + 1, D))[1]
        dp = _dp_scale(self.layer_drop, B * T, P + 1, x.device)
        qh, ph = self.attn.shadows()
        return ops.SpatialAttnFn.apply(
            x, _f32(self.norm.weight), _f32(self.norm.bias), _f32(self.attn.qkv.weight), _f32(self.attn.qkv.bias),
            _f32(self.attn.proj.weight), _f32(self.attn.proj.bias), qh, ph, dp, T, self.num_heads, self.norm.eps)


class MultiheadAttentionWithPreNorm(nn.Module):
    """Pre-norm joint self-attention with residual (ViViT encoders)."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0., norm_layer=nn.LayerNorm,
                 layer_drop=None, batch_first=False, **kwargs):
        super().__init__()
        if layer_drop is None:
            layer_drop = dict(type=DropPath, dropout_p=0.)
        self.embed_dims = embed_dims
        self.num_heads = num_heads
        self.norm = norm_layer(embed_dims)
        self.attn = Attention(embed_dims, num_heads, qkv_bias=True, attn_drop=attn_drop)
        self.proj_drop = nn.Dropout(proj_drop)
        self.layer_drop = _make_layer_drop(layer_drop)
        if proj_drop:
            raise NotImplementedError('proj_drop > 0 is not on the reference hot path')

    def forward(self, query, key=None, value=None, residual=None, attn_mask=None, key_padding_mask=None,
                return_attention=False, **kwargs):
        x = _f32(query).contiguous()
        Bp, N, D = x.shape
        if return_attention:
            xn = ops.RowsNormFn.apply(x, _f32(self.norm.weight), _f32(self.norm.bias), self.norm.eps, None)
            return self.attn(xn.view(Bp, N, D))[1]
        dp = _dp_scale(self.layer_drop, Bp, N, x.device)
        qh, ph = self.attn.shadows()
        return ops.JointAttnFn.apply(
            x, _f32(self.norm.weight), _f32(self.norm.bias), _f32(self.attn.qkv.weight), _f32(self.attn.qkv.bias),
            _f32(self.attn.proj.weight), _f32(self.attn.proj.bias), qh, ph, dp, self.num_heads, self.norm.eps)


class FFNWithPreNorm(nn.Module):
    """Pre-norm 2-layer MLP with exact-erf GELU and residual."""

    def __init__(self, embed_dims=256, hidden_channels=1024, num_layers=2, act_layer=nn.GELU,
                 norm_layer=nn.LayerNorm, dropout_p=0., layer_drop=None, **kwargs):
        super().__init__()
        assert num_layers >= 2, f'num_layers should be no less than 2. got {num_layers}.'
        if num_layers != 2 or act_layer is not nn.GELU or dropout_p:
            raise NotImplementedError('hot path covers the reference configuration: 2 layers, nn.GELU, dropout 0')
        self.embed_dims = embed_dims
        self.hidden_channels = hidden_channels
        self.num_layers = num_layers
        self.norm = norm_layer(embed_dims)
        layers = [nn.Sequential(nn.Linear(embed_dims, hidden_channels), act_layer(), nn.Dropout(dropout_p)),
                  nn.Linear(hidden_channels, embed_dims), nn.Dropout(dropout_p)]
        self.layers = nn.ModuleList(layers)
        self.layer_drop = _make_layer_drop(layer_drop)
        self._shadow = ShadowWeights()

    def forward(self, x):
        x = _f32(x).contiguous()
        fc1, fc2 = self.layers[0][0], self.layers[1]
        n0 = x.shape[0]
        dp = _dp_scale(self.layer_drop, n0, x.numel() // (n0 * x.shape[-1]), x.device)
        return ops.FFNFn.apply(x, _f32(self.norm.weight), _f32(self.norm.bias), _f32(fc1.weight), _f32(fc1.bias),
                               _f32(fc2.weight), _f32(fc2.bias), self._shadow.get('w1', fc1.weight),
                               self._shadow.get('w2', fc2.weight), dp, self.norm.eps)


class TransformerContainer(nn.Module):

    def __init__(self, num_transformer_layers, embed_dims, num_heads, num_frames, hidden_channels, operator_order,
                 drop_path_rate=0.1, norm_layer=nn.LayerNorm, act_layer=nn.GELU, num_layers=2):
        super().__init__()
        self.layers = nn.ModuleList([])
        self.num_transformer_layers = num_transformer_layers
        dpr = np.linspace(0, drop_path_rate, num_transformer_layers)
        for i in range(num_transformer_layers):
            self.layers.append(BasicTransformerBlock(
                embed_dims=embed_dims, num_heads=num_heads, num_frames=num_frames, hidden_channels=hidden_channels,
                operator_order=operator_order, norm_layer=norm_layer, act_layer=act_layer, num_layers=num_layers,
                dpr=dpr[i]))

    def forward(self, x, return_attention=False):
        last = self.num_transformer_layers - 1
        for idx, layer in enumerate(self.layers):
            x = layer(x, return_attention=True) if (idx >= last and return_attention) else layer(x)
        return x


class BasicTransformerBlock(nn.Module):

    def __init__(self, embed_dims, num_heads, num_frames, hidden_channels, operator_order, norm_layer=nn.LayerNorm,
                 act_layer=nn.GELU, num_layers=2, dpr=0):
        super().__init__()
        self.attentions = nn.ModuleList([])
        self.ffns = nn.ModuleList([])
        n_ops = len(operator_order)
        for i, op in enumerate(operator_order):
            drop = dict(type=DropPath, dropout_p=dpr)
            if op == 'self_attn':
                self.attentions.append(MultiheadAttentionWithPreNorm(
                    embed_dims=embed_dims, num_heads=num_heads, batch_first=True, norm_layer=nn.LayerNorm,
                    layer_drop=drop))
            elif op == 'time_attn':
                self.attentions.append(DividedTemporalAttentionWithPreNorm(
                    embed_dims=embed_dims, num_heads=num_heads, num_frames=num_frames, norm_layer=norm_layer,
                    use_cls_token=(i == n_ops - 2), layer_drop=drop))
            elif op == 'space_attn':
                self.attentions.append(DividedSpatialAttentionWithPreNorm(
                    embed_dims=embed_dims, num_heads=num_heads, num_frames=num_frames, norm_layer=norm_layer,
                    use_cls_token=(i == n_ops - 2), layer_drop=drop))
            elif op == 'ffn':
                self.ffns.append(FFNWithPreNorm(
                    embed_dims=embed_dims, hidden_channels=hidden_channels, num_layers=num_layers,
                    act_layer=act_layer, norm_layer=norm_layer, layer_drop=drop))
            else:
                raise TypeError(f'Unsupported operator type {op}')

    def forward(self, x, return_attention=False):
        n_attn = len(self.attentions)
        for idx, layer in enumerate(self.attentions):
            if idx >= n_attn - 1 and return_attention:
                return layer(x, return_attention=True)
            x = layer(x)
        for layer in self.ffns:
            x = layer(x)
        return x
