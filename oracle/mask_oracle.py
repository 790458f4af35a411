"""Restatement of the MaskFeat cube mask generator (reference mask_generator.py:23-107).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Consumes Python's `random` stream in
exactly the reference's order, so with the same seed the int32 mask and the
`[start, span]` markers are bit-identical (pinned by oracle/make_golden.py and the
SURVEY.md Appendix D known answers).
"""
from __future__ import annotations

import math
import random

import numpy as np


class CubeMaskOracle:
    def __init__(self, input_size=(8, 14, 14), mask_ratio=0.4, min_num_patches=16,
                 max_num_patches=None, min_aspect=0.3, max_aspect=None):
        self.T, self.H, self.W = input_size
        self.target_patches = int(self.H * self.W * mask_ratio)      # :30
        self.target_frames = int(self.T * mask_ratio)                # :31
        self.min_patches = min_num_patches
        self.max_patches = self.target_patches if max_num_patches is None else max_num_patches
        hi = max_aspect or 1 / min_aspect
        self.log_aspect = (math.log(min_aspect), math.log(hi))       # :36-37

    def _grow(self, m: np.ndarray, budget: int) -> int:
        """One block proposal round, mask_generator.py:48-70 (<=10 attempts)."""
        added = 0
        for _ in range(10):
            area = random.uniform(self.min_patches, budget)
            ar = math.exp(random.uniform(*self.log_aspect))
            h = int(round(math.sqrt(area * ar)))
            w = int(round(math.sqrt(area / ar)))
            if w < self.W and h < self.H:
                top = random.randint(0, self.H - h)
                left = random.randint(0, self.W - w)
                region = m[top:top + h, left:left + w]
                fresh = h * w - int(region.sum())
                if 0 < fresh <= budget:
                    region[...] = 1
                    added += fresh
            if added > 0:
                break
        return added

    def __call__(self):
        used = np.zeros(self.T, dtype=np.int32)
        cube = np.zeros((self.T, self.H, self.W), dtype=np.int32)
        markers = []
        done = 0
        while done < self.target_frames:                             # :77
            m = np.zeros((self.H, self.W), dtype=np.int32)
            count = 0
            while count < self.target_patches:                       # :81
                budget = min(self.target_patches - count, self.max_patches)
                d = self._grow(m, budget)
                if d == 0:
                    break
                count += d
            start = random.randint(0, self.T)                        # :91
            span = random.randint(1, self.target_frames - done)      # :92
            n = 0
            for i in range(start, start + span):
                if i > self.T - 1 or used[i]:
                    break
                used[i] = 1
                cube[i] = m
                n += 1
            done += n
            if n > 0:
                markers.append([start, n])
        return cube, markers


def center_frames(cube_marker, stride_t=2):
    """Centre-frame indices used by dataset.py:194 and video_transformer.py:894."""
    return [s * stride_t + n * stride_t // 2 for s, n in cube_marker]
