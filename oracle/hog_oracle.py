"""numpy fp64 restatement of the MaskFeat HOG target.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates ``extract_hog_features`` (reference dataset.py:39-45), i.e. three calls of
``skimage.feature.hog(channel, orientations=9, pixels_per_cell=(8,8),
cells_per_block=(1,1), block_norm='L2', feature_vector=False)`` followed by the
2x2 cell regroup.  The arithmetic lives in scikit-image (``feature/_hog.py`` and
``feature/_hoghistogram.pyx``; unpinned in reference requirements.txt:6, any
0.18/0.19 release is identical for this call).  scikit-image is NOT installed in
this image and the reference ships no HOG test vectors, so this restatement follows
the published algorithm and is **parity unpinned**.

Algorithm per colour channel (float64):
  1. un-normalised central differences, image borders zero   (_hog_channel_gradient)
  2. magnitude = hypot(gx, gy); orientation = rad2deg(arctan2(gy, gx)) % 180
  3. 8x8-pixel cells, 9 hard bins of 20 deg: hist[i,j,k] = sum(mag | 20k <= ori < 20(k+1)) / 64
     (cell_hog in _hoghistogram.pyx: `ori >= start or ori < end -> skip`)
  4. 1x1-cell blocks, L2: out = hist / sqrt(sum(hist^2) + 1e-5^2)  (_hog_normalize_block)
"""
from __future__ import annotations

import numpy as np

ORIENTATIONS = 9
CELL = 8
EPS = 1e-5


def channel_gradients(ch: np.ndarray):
    """(g_row, g_col) float64; first/last row (col) are zero."""
    img = ch.astype(np.float64)
    g_row = np.zeros_like(img)
    g_col = np.zeros_like(img)
    g_row[1:-1, :] = img[2:, :] - img[:-2, :]
    g_col[:, 1:-1] = img[:, 2:] - img[:, :-2]
    return g_row, g_col


def orientation_bins(g_row: np.ndarray, g_col: np.ndarray) -> np.ndarray:
    """Hard bin index 0..8; 9 = falls in no bin (ori rounds to exactly 180.0)."""
    ori = np.rad2deg(np.arctan2(g_row, g_col)) % 180
    width = 180.0 / ORIENTATIONS
    bins = np.full(ori.shape, ORIENTATIONS, dtype=np.uint8)
    for k in range(ORIENTATIONS):
        sel = (ori >= width * k) & (ori < width * (k + 1))
        bins[sel] = k
    return bins


def bin_lut() -> np.ndarray:
    """511x511 uint8 table: lut[gy+255, gx+255] = bin of integer gradient (gy, gx),
    built with the same numpy calls as `orientation_bins` (exact by construction)."""
    g = np.arange(-255, 256, dtype=np.float64)
    gy, gx = np.meshgrid(g, g, indexing='ij')
    return orientation_bins(gy, gx)


def hog_channel(ch: np.ndarray):
    """-> (normalised (Hc, Wc, 9) float64, bins (H, W) uint8)."""
    H, W = ch.shape
    g_row, g_col = channel_gradients(ch)
    mag = np.hypot(g_col, g_row)
    bins = orientation_bins(g_row, g_col)
    hc, wc = H // CELL, W // CELL
    hist = np.zeros((hc, wc, ORIENTATIONS), dtype=np.float64)
    magc = mag[:hc * CELL, :wc * CELL].reshape(hc, CELL, wc, CELL)
    binc = bins[:hc * CELL, :wc * CELL].reshape(hc, CELL, wc, CELL)
    for k in range(ORIENTATIONS):
        hist[:, :, k] = np.where(binc == k, magc, 0.0).sum(axis=(1, 3)) / (CELL * CELL)
    norm = np.sqrt((hist ** 2).sum(axis=-1, keepdims=True) + EPS ** 2)
    return hist / norm, bins


def extract_hog_features(image: np.ndarray) -> np.ndarray:
    """uint8 (H, W, 3) -> float64 (H/16, W/16, 108); feature index
    ((dh*2+dw)*27 + colour*9 + k), dataset.py:43-44."""
    H, W, _ = image.shape
    per = [hog_channel(image[:, :, c])[0] for c in range(3)]
    f = np.concatenate(per, axis=-1)                       # (Hc, Wc, 27)
    hc, wc = f.shape[:2]
    f = f.reshape(hc // 2, 2, wc // 2, 2, 27).transpose(0, 2, 1, 3, 4)
    return f.reshape(hc // 2, wc // 2, 108)


def extract_hog_bins(image: np.ndarray) -> np.ndarray:
    """uint8 (H, W, 3) -> uint8 (3, H, W) bin indices (9 = no bin)."""
    return np.stack([hog_channel(image[:, :, c])[1] for c in range(3)], axis=0)


def hog_targets(video_u8: np.ndarray, cube_marker) -> np.ndarray:
    """dataset.py:188-196: zeros (T,14,14,108) with HOG only on each cube's centre frame
    `start*2 + span*2//2`."""
    T, H, W, _ = video_u8.shape
    out = np.zeros((T, H // 16, W // 16, 108), dtype=np.float64)
    for start, span in cube_marker:
        c = start * 2 + span * 2 // 2
        out[c] = extract_hog_features(video_u8[c])
    return out
