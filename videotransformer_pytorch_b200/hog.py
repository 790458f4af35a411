"""MaskFeat HOG targets on the GPU (reference dataset.py:39-45, :188-196 -> skimage.feature.hog).

`hog_features(frames_u8)` runs the vt_hog kernel: integer gradients, orientation bins through a
511x511 look-up table that is built once on the host with numpy's own arctan2 / rad2deg / % 180 (so the
bin index of every possible integer gradient pair is bit-identical to the reference's float64 path),
fp32 magnitudes and cell sums, per-cell L2 normalisation, 2x2 cell regroup to (H/16, W/16, 108).
"""
from __future__ import annotations

import functools

import numpy as np
import torch

from . import _lib

ORIENTATIONS = 9


@functools.lru_cache(maxsize=1)
def _bin_lut_host() -> np.ndarray:
    g = np.arange(-255, 256, dtype=np.float64)
    gy, gx = np.meshgrid(g, g, indexing='ij')
    ori = np.rad2deg(np.arctan2(gy, gx)) % 180            # skimage _hoghistogram.pyx: orientation in [0, 180)
    lut = np.full(ori.shape, ORIENTATIONS, dtype=np.uint8)
    width = 180.0 / ORIENTATIONS
    for k in range(ORIENTATIONS):
        lut[(ori >= width * k) & (ori < width * (k + 1))] = k
    return np.ascontiguousarray(lut)


@functools.lru_cache(maxsize=8)
def _bin_lut(device: str) -> torch.Tensor:
    return torch.from_numpy(_bin_lut_host()).to(device).contiguous()


def hog_features(frames: torch.Tensor, want_bins: bool = False):
    """frames: uint8 [F, H, W, 3] on a CUDA device -> (fp32 [F, H/16, W/16, 108], uint8 [F,3,H,W] | None)."""
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise RuntimeError('hog_features: expected uint8 [F, H, W, 3]')
    return _lib.K.hog(frames, _bin_lut(str(frames.device)), want_bins)


def hog_targets(video_u8: torch.Tensor, cube_marker, dtype=torch.float32) -> torch.Tensor:
    """dataset.py:188-196 on device: zeros (T, H/16, W/16, 108) with HOG only on each mask cube's centre
    frame `start*2 + span*2//2`.  video_u8: uint8 [T, H, W, 3] (CUDA).  All centre frames of the clip go
    through one kernel launch."""
    T, H, W, _ = video_u8.shape
    out = torch.zeros((T, H // 16, W // 16, 108), dtype=dtype, device=video_u8.device)
    idx = [s * 2 + n * 2 // 2 for s, n in cube_marker]
    if idx:
        sel = torch.as_tensor(idx, device=video_u8.device, dtype=torch.long)
        feat, _ = hog_features(video_u8.index_select(0, sel).contiguous())
        out[sel] = feat.to(dtype)
    return out


def hog_targets_batch(video_u8: torch.Tensor, cube_markers, dtype=torch.float32) -> torch.Tensor:
    """hog_targets for a whole batch with ONE kernel launch: video_u8 uint8 [B, T, H, W, 3] (CUDA), cube_markers = one
    [[start, span], ...] list per sample -> [B, T, H/16, W/16, 108], zero except on every cube's centre frame."""
    B, T, H, W, _ = video_u8.shape
    out = torch.zeros((B, T, H // 16, W // 16, 108), dtype=dtype, device=video_u8.device)
    flat_idx = sorted({b * T + s * 2 + n * 2 // 2 for b, cm in enumerate(cube_markers) for s, n in cm})
    if flat_idx:
        sel = torch.as_tensor(flat_idx, device=video_u8.device, dtype=torch.long)
        frames = video_u8.reshape(B * T, H, W, 3).index_select(0, sel).contiguous()
        feat, _ = hog_features(frames)
        out.view(B * T, H // 16, W // 16, 108)[sel] = feat.to(dtype)
    return out
