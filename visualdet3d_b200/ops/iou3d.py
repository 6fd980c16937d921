"""`iou3d_cuda`-compatible entry points over libvd3d_b200 (R/lib/ops/iou3d).

Same signatures and in-place contracts as the pybind module (iou3d.cpp:174-179):
  boxes_overlap_bev_gpu(boxes_a[M,5], boxes_b[N,5], ans_overlap[M,N]) -> 1       (writes ans in place)
  boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou) -> 1
  nms_gpu(boxes[N,5] cuda (sorted by score), keep[N] int64 CPU tensor, thresh) -> num_to_keep   (fills keep[:num])
  nms_normal_gpu(...)  same with axis-aligned IoU
Errors: non-CUDA / non-contiguous inputs raise RuntimeError (the reference's CHECK_INPUT); nothing calls exit().
Unlike the reference the NMS sweep runs on the device and on the CURRENT stream (the reference uses the legacy default
stream and a synchronous cudaMemcpy); the only host sync is the final read of (count, keep).
Plus the Python-level helpers of iou3d.py:8-69 (`boxes3d_to_bev_torch`, `boxes_iou_bev`, `boxes_iou3d_gpu`).
The reference's own Python `nms_gpu` / `nms_normal_gpu` wrappers (iou3d.py:72-103) shadow the C symbols and recurse with the
wrong arguments; the C semantics are what is mirrored here.
"""
from __future__ import annotations

import torch

from .. import _lib
from .._lib import call


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _check(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap) -> int:
    for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_overlap, "ans_overlap")):
        _check(t, n)
    call("vd3d_boxes_overlap_bev", boxes_a.data_ptr(), boxes_a.shape[0], boxes_b.data_ptr(), boxes_b.shape[0], ans_overlap.data_ptr(), _stream())
    return 1


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou) -> int:
    for t, n in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_iou, "ans_iou")):
        _check(t, n)
    call("vd3d_boxes_iou_bev", boxes_a.data_ptr(), boxes_a.shape[0], boxes_b.data_ptr(), boxes_b.shape[0], ans_iou.data_ptr(), _stream())
    return 1


def _nms(boxes, keep, thresh: float, rotated: int) -> int:
    _check(boxes, "boxes")
    if not keep.is_contiguous() or keep.dtype != torch.int64:
        raise RuntimeError("keep must be a contiguous int64 tensor")
    N = boxes.shape[0]
    if N == 0:
        return 0
    dev = boxes.device
    ws = torch.empty(int(_lib.load().vd3d_nms_bev_workspace(N)), dtype=torch.uint8, device=dev)
    keep_dev = torch.empty(N, dtype=torch.int64, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    call("vd3d_nms_bev", boxes.data_ptr(), N, float(thresh), rotated, ws.data_ptr(), keep_dev.data_ptr(), count.data_ptr(), _stream())
    n = int(count.item())
    keep[:n].copy_(keep_dev[:n])
    return n


def nms_gpu(boxes, keep, nms_overlap_thresh: float) -> int:
    return _nms(boxes, keep, nms_overlap_thresh, 1)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh: float) -> int:
    return _nms(boxes, keep, nms_overlap_thresh, 0)


# ---- Python-level helpers (iou3d.py:8-69) -----------------------------------------------------------------------
def boxes3d_to_bev_torch(boxes3d):
    """(N, 7) [x, y, z, h, w, l, ry] -> (N, 5) [x1, y1, x2, y2, ry] in the x-z plane."""
    bev = boxes3d.new_empty((boxes3d.shape[0], 5))
    cu, cv = boxes3d[:, 0], boxes3d[:, 2]
    hl, hw = boxes3d[:, 5] / 2, boxes3d[:, 4] / 2
    bev[:, 0], bev[:, 1], bev[:, 2], bev[:, 3], bev[:, 4] = cu - hl, cv - hw, cu + hl, cv + hw, boxes3d[:, 6]
    return bev


def boxes_iou_bev(boxes_a, boxes_b):
    ans = torch.zeros(boxes_a.shape[0], boxes_b.shape[0], dtype=torch.float32, device=boxes_a.device)
    boxes_iou_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans)
    return ans


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """(N,7),(M,7) [x, y, z, h, w, l, ry] -> 3-D IoU (N, M): BEV overlap x height overlap / union volume (iou3d.py:37-69)."""
    a_bev, b_bev = boxes3d_to_bev_torch(boxes_a), boxes3d_to_bev_torch(boxes_b)
    ov = torch.zeros(boxes_a.shape[0], boxes_b.shape[0], dtype=torch.float32, device=boxes_a.device)
    boxes_overlap_bev_gpu(a_bev.contiguous(), b_bev.contiguous(), ov)
    a_min, a_max = (boxes_a[:, 1] - boxes_a[:, 3]).view(-1, 1), boxes_a[:, 1].view(-1, 1)
    b_min, b_max = (boxes_b[:, 1] - boxes_b[:, 3]).view(1, -1), boxes_b[:, 1].view(1, -1)
    oh = torch.clamp(torch.min(a_max, b_max) - torch.max(a_min, b_min), min=0)
    o3 = ov * oh
    va = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vb = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    return o3 / torch.clamp(va + vb - o3, min=1e-7)
