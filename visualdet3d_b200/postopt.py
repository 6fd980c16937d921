"""`post_optimization` of the anchor-based 3-D heads: hill climbing on the yaw so that the projected
3-D box matches the detected 2-D box (R/heads/detection_3d_head.py:294-308 -> R/lib/fast_utils/hill_climbing.py:7-122).

The search itself is `vd3d_post_opt_host` of the C-ABI library (float64, one routine shared with the CUDA kernel `vd3d_post_opt`);
this module is the host glue of `post_opt`: which rows are refined (label 0 and z > 3 m), the float32 alpha <-> yaw conversions
done with the very numpy calls the reference makes (so their float32 rounding is the reference's), and the float32 write-back.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib

IMG_W, IMG_H = 1280.0, 288.0          # the reference clips the projected hull to these constants (hill_climbing.py:98-103)


def _alpha_to_rot(alpha: np.ndarray, cx: float, P2: np.ndarray) -> np.ndarray:
    """visualDet3D/utils/utils.py:30-36 on a one-element array (float32 in, float32 out under NumPy >= 2 promotion rules)."""
    ry = alpha + np.arctan2(cx - P2[..., 0, 2], P2[..., 0, 0])
    ry[np.where(ry > np.pi)] -= 2 * np.pi
    ry[np.where(ry <= -np.pi)] += 2 * np.pi
    return ry


def _rot_to_alpha(ry: np.ndarray, cx: float, P2: np.ndarray) -> np.ndarray:
    """visualDet3D/utils/utils.py:39-45"""
    alpha = ry - np.arctan2(cx - P2[..., 0, 2], P2[..., 0, 0])
    alpha[alpha > np.pi] -= 2 * np.pi
    alpha[alpha <= -np.pi] += 2 * np.pi
    return alpha


def post_process(bboxes: torch.Tensor, labels: torch.Tensor, P2, step_r_init: float = 0.4, r_lim: float = 0.01,
                 min_depth: float = 3.0, label: int = 0, img_w: float = IMG_W, img_h: float = IMG_H) -> torch.Tensor:
    """bboxes [K, 11] = (x1, y1, x2, y2, cx, cy, z, w, h, l, alpha) on the host, labels [K], P2 [3, 4] -> refined copy.
    Rows with `labels == label` and `z > min_depth` get their alpha re-estimated; everything else is returned unchanged."""
    out = bboxes.detach().cpu().float().clone()
    if out.shape[0] == 0:
        return out
    P2 = np.asarray(P2.detach().cpu().numpy() if isinstance(P2, torch.Tensor) else P2)
    lab = labels.detach().cpu().numpy()
    box = out.numpy()
    sel = [i for i in range(box.shape[0]) if box[i, 6] > min_depth and lab[i] == label]
    if not sel:
        return out
    p2 = np.eye(4)
    p2[0:3] = P2.copy()
    p2_inv = np.ascontiguousarray(np.linalg.inv(p2))
    n = len(sel)
    cx = np.array([float(box[i, 4]) for i in sel], dtype=np.float64)            # `.item()` of the float32 tensor element
    cy = np.array([float(box[i, 5]) for i in sel], dtype=np.float64)
    theta0 = np.array([_alpha_to_rot(np.array([box[i, 10]]), float(box[i, 4]), P2)[0] for i in sel], dtype=np.float32)
    b2 = np.ascontiguousarray(box[sel, 0:4], dtype=np.float32)
    z, w, h, l = (np.ascontiguousarray(box[sel, c], dtype=np.float32) for c in (6, 7, 8, 9))
    theta = np.empty(n, dtype=np.float64)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.call("vd3d_post_opt_host", ptr(np.ascontiguousarray(p2)), ptr(p2_inv), n, ptr(b2), ptr(cx), ptr(cy), ptr(z), ptr(w), ptr(h), ptr(l),
              ptr(theta0), float(img_w), float(img_h), float(step_r_init), float(r_lim), ptr(theta), None)
    for j, i in enumerate(sel):
        box[i, 10] = _rot_to_alpha(np.array([theta[j]]), float(box[i, 4]), P2)[0]       # float64 -> stored as float32, like `bbox3d_state_3d.new([...])`
    return out
