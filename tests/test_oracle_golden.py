"""The oracle restatements against the reference-generated golden vectors (CPU, no reference needed)."""
import hashlib
import random

import numpy as np
import pytest
import torch

from tests.conftest import check_grads, rel_err


@pytest.mark.parametrize('name', ['timesformer_tiny', 'timesformer_hd64'])
def test_timesformer_oracle_vs_golden(golden, name):
    from oracle import vt_oracle as O
    g = golden(name)
    sd = {k: v.double() for k, v in g.sd.items()}
    x = g.x.double()
    with torch.no_grad():
        assert rel_err(O.timesformer_forward(sd, x, g.cfg), g.out['y_eval']) < 1e-12
        assert rel_err(O.timesformer_tokens(sd, x, g.cfg), g.out['tokens']) < 1e-12
        assert rel_err(O.timesformer_last_selfattention(sd, x, g.cfg), g.out['last_attn']) < 1e-12
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xg = x.clone().requires_grad_(True)
    torch.manual_seed(g.train_seed)
    y = O.timesformer_forward(sdg, xg, g.cfg, training=True)
    assert rel_err(y, g.out['y_train']) < 1e-12
    (y * g.out['loss_w']).sum().backward()
    assert rel_err(xg.grad, g.out['dx']) < 1e-6          # golden dx stored as fp32
    check_grads({k: v.grad for k, v in sdg.items()}, g, 1e-6)


@pytest.mark.parametrize('name', ['vivit_tiny_b1', 'vivit_tiny_b3'])
def test_vivit_oracle_vs_golden(golden, name):
    from oracle import vt_oracle as O
    g = golden(name)
    sd = {k: v.double() for k, v in g.sd.items()}
    x = g.x.double()
    with torch.no_grad():
        assert rel_err(O.vivit_forward(sd, x, g.cfg), g.out['y_eval']) < 1e-12
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    torch.manual_seed(g.train_seed)
    y = O.vivit_forward(sdg, x, g.cfg, training=True)
    assert rel_err(y, g.out['y_train']) < 1e-12
    (y * g.out['loss_w']).sum().backward()
    check_grads({k: v.grad for k, v in sdg.items()}, g, 1e-6)


def test_vivit_cls_gather_quirk_is_batch_dependent(golden):
    """video_transformer.py:515 takes rows 0..B-1 of the (b t)-major tensor: sample j>0 sees sample 0's frames."""
    from oracle import vt_oracle as O
    g = golden('vivit_tiny_b3')
    sd = {k: v.double() for k, v in g.sd.items()}
    x = g.x.double()
    with torch.no_grad():
        full = O.vivit_forward(sd, x, g.cfg)
        alone = O.vivit_forward(sd, x[1:2], g.cfg)
    assert rel_err(full[0:1], O.vivit_forward(sd, x[0:1], g.cfg)) < 1e-12    # sample 0 is self-consistent
    assert rel_err(full[1:2], alone) > 1e-3                                     # sample 1 is not


def test_mask_generators_vs_golden():
    import os
    from oracle.mask_oracle import CubeMaskOracle
    from tests.conftest import GOLD
    from videotransformer_pytorch_b200.mask_generator import CubeMaskGenerator
    z = np.load(os.path.join(GOLD, 'cube_mask.npz'))
    for cls in (CubeMaskOracle, CubeMaskGenerator):
        for seed in range(8):
            random.seed(seed); np.random.seed(seed)
            gen = cls(input_size=(8, 14, 14), min_num_patches=16)
            flat = []
            for i in range(3):
                m, mk = gen()
                assert m.dtype == np.int32 and np.array_equal(m, z[f'mask_{seed}'][i]), (cls.__name__, seed, i)
                flat += [[i, s, n] for s, n in mk]
            assert np.array_equal(np.asarray(flat, dtype=np.int32).reshape(-1, 3), z[f'markers_{seed}'].reshape(-1, 3))
    # SURVEY.md Appendix D known answers
    assert hashlib.sha256(z['mask_0'][0].tobytes()).hexdigest()[:16] == 'c87b9c69a35b57eb'
    assert z['mask_0'][0].sum() == 234 and z['mask_0'][0].sum(axis=(1, 2)).tolist() == [78, 0, 78, 78, 0, 0, 0, 0]


def test_hog_oracle_properties():
    """No reference vectors exist for HOG (parity unpinned); check the restatement's defining properties."""
    from oracle import hog_oracle as HO
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(32, 48, 3), dtype=np.uint8)
    f = HO.extract_hog_features(img)
    assert f.shape == (2, 3, 108) and f.dtype == np.float64
    # per (cell, channel) 9-vectors have unit L2 norm (up to the 1e-5 eps) or are zero
    v = f.reshape(2, 3, 4, 3, 9)
    n = np.sqrt((v ** 2).sum(-1))
    assert np.all((np.abs(n - 1) < 1e-6) | (n < 1e-6))
    # brute-force per-pixel loop == vectorised implementation
    ch = img[:, :, 1].astype(np.float64)
    gr, gc = HO.channel_gradients(img[:, :, 1])
    hist = np.zeros((4, 6, 9))
    for y in range(32):
        for x in range(48):
            ori = np.rad2deg(np.arctan2(gr[y, x], gc[y, x])) % 180
            k = int(ori // 20)
            if k < 9:
                hist[y // 8, x // 8, k] += np.hypot(gc[y, x], gr[y, x])
    hist /= 64
    ref = hist / np.sqrt((hist ** 2).sum(-1, keepdims=True) + 1e-10)
    got, bins = HO.hog_channel(img[:, :, 1])
    assert np.allclose(got, ref, atol=1e-12)
    # LUT == direct binning for every gradient that occurs
    lut = HO.bin_lut()
    assert np.array_equal(lut[(gr.astype(int) + 255), (gc.astype(int) + 255)], bins)
    # horizontal ramp: gradient purely along x -> orientation 0 -> bin 0 everywhere inside
    ramp = np.tile(np.arange(48, dtype=np.uint8) * 4, (32, 1))
    _, b = HO.hog_channel(ramp)
    assert np.all(b[1:-1, 1:-1] == 0)
    # the product's LUT is the same table
    from videotransformer_pytorch_b200.hog import _bin_lut_host
    assert np.array_equal(_bin_lut_host(), lut)


def test_space_only_oracle_vs_golden(golden):
    from oracle import vt_oracle as O
    g = golden('timesformer_space_only_tiny')
    sd = {k: v.double() for k, v in g.sd.items()}
    x = g.x.double()
    with torch.no_grad():
        assert rel_err(O.timesformer_space_only_forward(sd, x, g.cfg), g.out['y_eval']) < 1e-12
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    torch.manual_seed(g.train_seed)
    y = O.timesformer_space_only_forward(sdg, x, g.cfg, training=True)
    assert rel_err(y, g.out['y_train']) < 1e-12
    (y * g.out['loss_w']).sum().backward()
    check_grads({k: v.grad for k, v in sdg.items()}, g, 1e-6)


@pytest.mark.parametrize('name', ['timesformer_joint_tiny', 'timesformer_joint_n289'])
def test_timesformer_joint_oracle_vs_golden(golden, name):
    from oracle import vt_oracle as O
    g = golden(name)
    sd = {k: v.double() for k, v in g.sd.items()}
    x = g.x.double()
    with torch.no_grad():
        assert rel_err(O.timesformer_joint_forward(sd, x, g.cfg), g.out['y_eval']) < 1e-12
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    torch.manual_seed(g.train_seed)
    y = O.timesformer_joint_forward(sdg, x, g.cfg, training=True)
    assert rel_err(y, g.out['y_train']) < 1e-12
    (y * g.out['loss_w']).sum().backward()
    check_grads({k: v.grad for k, v in sdg.items()}, g, 1e-6)
