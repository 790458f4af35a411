"""MaskFeat cube-mask generation (reference mask_generator.py:23-107) — host-side integer logic.

This is a parity item, not a GPU kernel (224 us/call of Python `random` draws whose
order defines the result).  The implementation consumes the `random` stream in the reference's order,
so for equal seeds the int32 mask and the [start, span] markers are bit-identical.
"""
from __future__ import annotations

import math
import random

import numpy as np


class CubeMaskGenerator:
    def __init__(self, input_size=(8, 14, 14), mask_ratio=0.4, min_num_patches=16, max_num_patches=None,
                 min_aspect=0.3, max_aspect=None):
        self.temporal, self.height, self.width = input_size
        self.num_patches = self.height * self.width
        self.num_masking_patches = int(self.num_patches * mask_ratio)
        self.num_masking_frames = int(self.temporal * mask_ratio)
        self.min_num_patches = min_num_patches
        self.max_num_patches = self.num_masking_patches if max_num_patches is None else max_num_patches
        max_aspect = max_aspect or 1 / min_aspect
        self.log_aspect_ratio = (math.log(min_aspect), math.log(max_aspect))

    def __repr__(self):
        return 'Generator(%d, %d -> [%d ~ %d], max = %d, %.3f ~ %.3f)' % (
            self.height, self.width, self.min_num_patches, self.max_num_patches, self.num_masking_patches,
            self.log_aspect_ratio[0], self.log_aspect_ratio[1])

    def get_shape(self):
        return self.temporal, self.height, self.width

    def _mask(self, mask, max_mask_patches):
        """Up to 10 block proposals; returns the number of newly masked patches."""
        gained = 0
        for _ in range(10):
            area = random.uniform(self.min_num_patches, max_mask_patches)
            aspect = math.exp(random.uniform(*self.log_aspect_ratio))
            h = int(round(math.sqrt(area * aspect)))
            w = int(round(math.sqrt(area / aspect)))
            if w < self.width and h < self.height:
                top = random.randint(0, self.height - h)
                left = random.randint(0, self.width - w)
                block = mask[top:top + h, left:left + w]
                new = h * w - int(block.sum())
                if 0 < new <= max_mask_patches:
                    block[...] = 1
                    gained += new
            if gained > 0:
                break
        return gained

    def __call__(self):
        taken = np.zeros(self.temporal, dtype=np.int32)
        cube_mask = np.zeros(self.get_shape(), dtype=np.int32)
        cube_marker = []
        frames_done = 0
        while frames_done < self.num_masking_frames:
            plane = np.zeros((self.height, self.width), dtype=np.int32)
            covered = 0
            while covered < self.num_masking_patches:
                budget = min(self.num_masking_patches - covered, self.max_num_patches)
                got = self._mask(plane, budget)
                if got == 0:
                    break
                covered += got
            start = random.randint(0, self.temporal)
            span = random.randint(1, self.num_masking_frames - frames_done)
            placed = 0
            for t in range(start, start + span):
                if t > self.temporal - 1 or taken[t]:
                    break
                taken[t] = 1
                cube_mask[t] = plane
                placed += 1
            frames_done += placed
            if placed > 0:
                cube_marker.append([start, placed])
        return cube_mask, cube_marker


class RandomMaskGenerator:
    """reference mask_generator.py:5-21."""

    def __init__(self, input_size=224, mask_ratio=0.6):
        if not isinstance(input_size, tuple):
            input_size = (input_size,) * 2
        self.height, self.width = input_size
        self.num_patches = self.height * self.width
        self.num_mask = int(mask_ratio * self.num_patches)

    def __call__(self):
        mask = np.hstack([np.zeros(self.num_patches - self.num_mask), np.ones(self.num_mask)])
        np.random.shuffle(mask)
        return mask
