"""Host-side anchor table + 3-D priors (the reference builds these on the host in float64 numpy too:
R/heads/anchors.py:59-91, generate_anchors :152-183, shift :219-239, anchors2indexes :45-57).

The table is a function of the image shape and the config only, so it is built once per shape and uploaded; the
per-frame part (the P2-dependent useful mask, anchors.py:93-111) is a CUDA kernel (csrc/postprocess.cu).

Index contract (must match AnchorFlatten, R/lib/blocks.py:134-135): n = (y * Wf + x) * A + a with
a = ratio_idx * n_scales + scale_idx, i.e. conv output channel = a * C_out + c.
"""
from __future__ import annotations

import os
from typing import Sequence, Tuple

import numpy as np
import torch


def load_priors(preprocessed_path: str, obj_types: Sequence[str], n_rows: int, n_ratios: int, channels: int = 6):
    """anchor_{mean,std}_{type}.npy under {preprocessed_path}/training (anchors.py:30-40) -> two [T, rows, ratios, 6] f64."""
    mean = np.zeros([len(obj_types), n_rows, n_ratios, channels])
    std = np.zeros_like(mean)
    d = os.path.join(preprocessed_path, "training")
    for i, t in enumerate(obj_types):
        mean[i] = np.load(os.path.join(d, f"anchor_mean_{t}.npy"))
        std[i] = np.load(os.path.join(d, f"anchor_std_{t}.npy"))
    return mean, std


def cell_anchors(base_size: float, ratios, scales) -> np.ndarray:
    """The A = len(ratios)*len(scales) zero-centred boxes of one cell, float64, ratio-major order."""
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    r = np.repeat(ratios, scales.size)          # ratio of anchor a
    s = np.tile(scales, ratios.size)            # scale of anchor a
    edge = base_size * s
    area = edge * edge
    w = np.sqrt(area / r)
    h = w * r
    half_w, half_h = w * 0.5, h * 0.5
    return np.stack([0.0 - half_w, 0.0 - half_h, w - half_w, h - half_h], axis=1)


def grid_anchors(image_hw: Sequence[int], levels, strides, sizes, ratios, scales) -> np.ndarray:
    """All anchors of all pyramid levels, float64 [N, 4], x fastest then y then per-cell index."""
    hw = np.asarray(image_hw, dtype=np.int64)
    chunks = []
    for lv, stride, size in zip(levels, strides, sizes):
        fh, fw = (hw + 2 ** lv - 1) // (2 ** lv)            # ceil division (anchors.py:66)
        cx = (np.arange(fw) + 0.5) * stride
        cy = (np.arange(fh) + 0.5) * stride
        gx, gy = np.meshgrid(cx, cy)                         # [fh, fw]
        centres = np.stack([gx, gy, gx, gy], axis=-1).reshape(-1, 1, 4)
        chunks.append((centres + cell_anchors(size, ratios, scales)[None]).reshape(-1, 4))
    return np.concatenate(chunks, axis=0)


def prior_lookup(anchors64: np.ndarray, sizes, ratios, scales):
    """(scale row, ratio column) of every anchor in the prior tables: nearest size / nearest ratio (anchors.py:45-57)."""
    w = anchors64[:, 2] - anchors64[:, 0]
    h = anchors64[:, 3] - anchors64[:, 1]
    size_grid = (np.asarray(sizes, dtype=np.float64) * np.asarray(scales, dtype=np.float64))[:, None]
    row = np.argmin(np.abs(np.sqrt(w * h) - size_grid), axis=0)
    col = np.argmin(np.abs(h / w - np.asarray(ratios, dtype=np.float64)[:, None]), axis=0)
    return row, col


class AnchorTable:
    """Device-resident anchors [N,4] f32, priors mean_std [N,T,6,2] f32 and z-means [T,N] f32 for one image shape."""

    def __init__(self, image_hw, anchors_cfg, prior_mean: np.ndarray, prior_std: np.ndarray, device):
        a64 = grid_anchors(image_hw, anchors_cfg["pyramid_levels"], anchors_cfg["strides"], anchors_cfg["sizes"],
                           anchors_cfg["ratios"], anchors_cfg["scales"])
        row, col = prior_lookup(a64, anchors_cfg["sizes"], anchors_cfg["ratios"], anchors_cfg["scales"])
        means = torch.tensor(prior_mean[:, row, col], dtype=torch.float32)       # [T, N, 6]
        stds = torch.tensor(prior_std[:, row, col], dtype=torch.float32)
        self.image_hw = tuple(int(v) for v in image_hw)
        self.N = a64.shape[0]
        self.T = means.shape[0]
        self.anchors = torch.tensor(a64.astype(np.float32)).to(device).contiguous()
        self.mean_std = torch.stack([means, stds], dim=-1).permute(1, 0, 2, 3).contiguous().to(device)   # [N,T,6,2]
        self.means_z = means[:, :, 0].contiguous().to(device)                                         # [T,N]
        self.num_anchors_per_cell = len(anchors_cfg["ratios"]) * len(anchors_cfg["scales"]) * len(anchors_cfg["pyramid_levels"])
