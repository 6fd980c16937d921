"""Post-forward geometry and the KITTI result writer: what `test_one`
(R/pipelines/evaluators.py:101-145) does between the detector's output and the result file.

    scores[K], bbox[K, 11] = (x1, y1, x2, y2, cx, cy, z, w, h, l, alpha), cls[K]
      -> back-projection of (cx, cy, z) to camera coordinates      (BackProjection,  R/utils/utils.py:256-278)
      -> observation angle alpha -> rotation theta (+ box corners)  (BBox3dProjector, R/utils/utils.py:198-254; alpha2theta_3d, visualDet3D/utils/utils.py:47-62)
      -> 2-D boxes moved / scaled back to the original image        (evaluators.py:118-127; 2-D-only branch :129-143)
      -> one KITTI label line per detection above the threshold     (write_result_to_file, data/kitti/utils.py:162-201)

Everything here runs on the HOST, on the detection records that `StreamedInference.collect` / `forward_batch(...).cpu()` already
brought back (K <= 512 rows per image): it costs no GPU time and no extra synchronisation (the reference runs the same handful of
elementwise ops as ~20 eager CUDA launches plus `.item()` reads per frame).  The arithmetic is float32 in the reference's operation
order, so the formatted lines are character-identical to the reference's on the same inputs (tests/test_postforward_cpu.py against
fixtures generated from the unmodified reference).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

# corner signs of a box in its own frame (x: width, y: height, z: length), the order of BBox3dProjector.corner_matrix
_CORNERS = torch.tensor([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [1, 1, 1], [1, -1, 1], [-1, -1, 1], [-1, 1, 1], [-1, 1, -1]], dtype=torch.float32)


def _f32(x) -> torch.Tensor:
    return torch.as_tensor(x, dtype=torch.float32)


def back_projection(state: torch.Tensor, P2) -> torch.Tensor:
    """[K, >=7] rows (u, v, z, w, h, l, alpha, ...) with (u, v) the projected centre in pixels -> (x, y, z, w, h, l, alpha, ...)
    in the camera frame of P2 [3, 4]:  x = (u z - cx z - tx) / fx,  y = (v z - cy z - ty) / fy."""
    s, P = _f32(state), _f32(P2)
    z = s[:, 2:3]
    x = (s[:, 0:1] * z - P[0, 2] * z - P[0, 3]) / P[0, 0]
    y = (s[:, 1:2] * z - P[1, 2] * z - P[1, 3]) / P[1, 1]
    return torch.cat([x, y, s[:, 2:]], dim=1)


def alpha_to_theta(alpha: torch.Tensor, x: torch.Tensor, z: torch.Tensor, P2) -> torch.Tensor:
    """theta = alpha + atan2(x + tx / fx, z)"""
    P = _f32(P2)
    return _f32(alpha) + torch.atan2(_f32(x) + P[0, 3] / P[0, 0], _f32(z))


def project_boxes(box3d: torch.Tensor, P2) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """[K, >=7] camera-frame boxes (x, y, z, w, h, l, alpha) -> (corners in the camera frame [K, 8, 3],
    corners in the image (homogeneous, divided by depth + 1e-6) [K, 8, 3], theta [K])."""
    b, P = _f32(box3d), _f32(P2)
    rel = 0.5 * _CORNERS * b[:, 3:6].unsqueeze(1)                       # [K, 8, 3]
    theta = alpha_to_theta(b[:, 6], b[:, 0], b[:, 2], P)
    c, s = torch.cos(theta).unsqueeze(1), torch.sin(theta).unsqueeze(1)
    rx = rel[:, :, 2] * c + rel[:, :, 0] * s
    rz = -rel[:, :, 2] * s + rel[:, :, 0] * c
    corners = torch.stack([rx, rel[:, :, 1], rz], dim=-1) + b[:, 0:3].unsqueeze(1)
    hom = torch.cat([corners, corners.new_ones(corners.shape[0], 8, 1)], dim=-1).unsqueeze(3)      # [K, 8, 4, 1]
    cam = torch.matmul(P, hom).squeeze(-1)
    return corners, cam / (cam[:, :, 2:] + 1e-6), theta


def rescale_boxes_2d(box2d: torch.Tensor, P2, original_P) -> torch.Tensor:
    """2-D boxes in network-input pixels -> pixels of the original image (undo crop + resize through the two calibrations)."""
    b = _f32(box2d).clone()
    P, O = np.asarray(P2), np.asarray(original_P)
    sx, sy = O[0, 0] / P[0, 0], O[1, 1] / P[1, 1]
    b[:, 0:4:2] += O[0, 2] / sx - P[0, 2]
    b[:, 1:4:2] += O[1, 2] / sy - P[1, 2]
    b[:, 0:4:2] *= sx
    b[:, 1:4:2] *= sy
    return b


def rescale_boxes_2d_only(box2d: torch.Tensor, net_height: int, original_height: int, crop_top: int) -> torch.Tensor:
    """2-D-only detectors (evaluators.py:129-143): uniform scale back to the cropped original, then the crop offset."""
    b = _f32(box2d).clone()
    b[:, 0:4] *= (original_height - crop_top) / net_height
    b[:, 1:4:2] += crop_top
    return b


def kitti_lines(scores, box2d, box3d=None, thetas=None, obj_types: Sequence[str] = ("Car", "Pedestrian", "Cyclist"), threshold: float = 0.4) -> str:
    """The text of one KITTI result file.  box3d rows = (x, y_centre, z, w, h, l, alpha): y is moved to the box bottom
    (`y + h / 2`, KITTI convention) on a copy; without 3-D boxes the reference's placeholders (-1 / -1000 / -10) are written."""
    k = len(box2d)
    if box3d is None:
        b3 = np.ones((k, 7), dtype=int)
        b3[:, 3:6], b3[:, 0:3], b3[:, 6] = -1, -1000, -10
    else:
        b3 = _f32(box3d).clone()
        b3[:, 1] = b3[:, 1] + 0.5 * b3[:, 4]
    th = np.ones(k) * -10 if thetas is None else thetas
    out = []
    if len(scores) > 0:
        # rows as Python floats once (formatting 0-dim tensors one by one costs ~10 us each); `{:.6f}` of a float32 tensor / numpy
        # scalar is the format of its exact double value, so the text does not change.  The trailing `{}` score keeps the element
        # type of `scores` (the reference prints a tensor element as the repr of its double value, a numpy float32 by its shortest repr).
        b2l = box2d.tolist() if hasattr(box2d, "tolist") else [list(map(float, r)) for r in box2d]
        b3l = b3.tolist()
        thl = th.tolist() if hasattr(th, "tolist") else list(th)
        keep = (scores >= threshold).tolist() if hasattr(scores, "tolist") else [sc >= threshold for sc in scores]
        for i in range(k):
            if not keep[i]:
                continue
            bb, s3 = b2l[i], b3l[i]
            out.append("{} -1 -1 {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {:.6f} {} \n".format(
                obj_types[i], s3[-1], bb[0], bb[1], bb[2], bb[3], s3[4], s3[3], s3[5], s3[0], s3[1], s3[2], thl[i], scores[i]))
    return "".join(out)


def detections_to_kitti(scores: torch.Tensor, bbox: torch.Tensor, cls: torch.Tensor, P2, original_P, class_names: Sequence[str],
                        threshold: float = 0.4) -> str:
    """Forward output of one image -> its KITTI result text (the 3-D branch of `test_one`)."""
    scores, bbox = _f32(scores).cpu(), _f32(bbox).cpu()
    names = [class_names[int(i)] for i in cls]
    box3d = back_projection(bbox[:, 4:], P2)
    theta = alpha_to_theta(box3d[:, 6], box3d[:, 0], box3d[:, 2], P2)
    box2d = rescale_boxes_2d(bbox[:, 0:4], P2, original_P)
    return kitti_lines(scores, box2d, box3d, theta, names, threshold)


def write_result_file(result_dir: str, index: int, text: str) -> str:
    path = os.path.join(result_dir, "%06d.txt" % index)
    with open(path, "w") as f:
        f.write(text)
    return path
