"""Attention core kernels vs torch fp32 on the same bf16 qkv.  -m gpu"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def K():
    from videotransformer_pytorch_b200 import _lib
    return _lib.K


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def ref_attn(qkv, Bp, N, H, hd, scale):
    q = qkv.float().reshape(Bp, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    s = (q[0] @ q[1].transpose(-1, -2)) * scale
    p = s.softmax(-1)
    o = (p @ q[2]).transpose(1, 2).reshape(Bp * N, H * hd)
    return o, p, torch.logsumexp(s, -1)


@pytest.mark.parametrize('Bp,N,H', [(6, 8, 2), (3, 9, 12), (4, 33, 2), (5, 197, 3), (2, 256, 1), (1568, 8, 12)])
def test_attn_fwd_bwd(Bp, N, H):
    hd = 64
    torch.manual_seed(N)
    qkv = (torch.randn(Bp, N, 3, H, hd) * 0.7).cuda().bfloat16()
    scale = hd ** -0.5
    ctx, lse, probs = K().attn_fwd(qkv, Bp, N, H, hd, scale, want_probs=True)
    qf = qkv.float().requires_grad_(True)
    o, p, l = ref_attn(qf, Bp, N, H, hd, scale)
    assert rel(ctx, o) < 4e-3, rel(ctx, o)
    assert rel(probs, p) < 1e-4
    assert rel(lse, l) < 1e-5
    dctx = torch.randn(Bp * N, H * hd).cuda().bfloat16()
    o.backward(dctx.float())
    dqkv = K().attn_bwd(qkv, ctx, dctx, lse, Bp, N, H, hd, scale)
    assert dqkv.shape == qkv.shape
    for i, nm in enumerate('qkv'):
        e = rel(dqkv[:, :, i], qf.grad[:, :, i])
        assert e < 1e-2, (nm, e)
