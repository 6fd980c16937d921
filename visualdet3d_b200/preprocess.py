"""Test-time input pipeline: what the reference's `test_augmentation` does per frame on the CPU with cv2 /
numpy (R/data/pipeline/stereo_augmentator.py: ConvertToFloat :29-36, CropTop :213-258, Resize :63-134, Normalize :39-60) and
`KittiStereoDataset.__getitem__` / `collate_fn` (R/data/kitti/dataset/stereo_dataset.py:176-203): uint8 HWC frame -> float32 CHW network
input, and the calibration matrices moved along.

`preprocess_host` / `preprocess_batch` call the C-ABI library (one routine shared by the host entry and the CUDA kernel); the calibration
update is host float arithmetic in the reference's order."""
from __future__ import annotations

import ctypes
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib

RGB_MEAN = (0.485, 0.456, 0.406)
RGB_STD = (0.229, 0.224, 0.225)


def adjust_calib(P: np.ndarray, crop_top: int, height: int, out_height: int) -> np.ndarray:
    """P [3, 4] of the original frame -> P of the network input: CropTop (cy -= dv, ty -= dv * tz) then Resize (rows 0 and 1 scaled)."""
    P = np.array(P, copy=True)
    P[1, 2] = P[1, 2] - crop_top
    P[1, 3] = P[1, 3] - crop_top * P[2, 3]
    s = out_height / (height - crop_top)
    P[0, :] = P[0, :] * s
    P[1, :] = P[1, :] * s
    return P


def _f32(v: Sequence[float]) -> np.ndarray:
    return np.ascontiguousarray(np.array(v, dtype=np.float32))


def preprocess_host(frame: np.ndarray, crop_top: int, size: Tuple[int, int], mean=RGB_MEAN, std=RGB_STD) -> np.ndarray:
    """uint8 [H, W, C] -> float32 [C, size[0], size[1]] on the host: the parity checker of the CUDA form (`preprocess_batch` is the product
    path; tests pin this routine to the reference and the kernel to this routine)."""
    assert frame.dtype == np.uint8 and frame.ndim == 3
    frame = np.ascontiguousarray(frame)
    H, W, C = frame.shape
    out = np.empty((C, size[0], size[1]), dtype=np.float32)
    m, s = _f32(mean), _f32(std)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.call("vd3d_preprocess_host", vp(frame), H, W, C, W * C, int(crop_top), int(size[0]), int(size[1]), vp(m), vp(s), vp(out))
    return out


def preprocess_batch(frames: List[np.ndarray], crop_top: int, size: Tuple[int, int], mean=RGB_MEAN, std=RGB_STD, device="cuda") -> torch.Tensor:
    """uint8 HWC frames (sizes may differ) -> [B, C, size[0], size[1]] float32 on `device`: one uint8 upload per frame (3 bytes per pixel
    instead of 12) and one kernel for the whole batch."""
    lib = _lib.load()
    B = len(frames)
    C = frames[0].shape[2]
    nb = int(lib.vd3d_preprocess_desc_bytes())
    descs = np.zeros((B, nb), dtype=np.uint8)
    dev_frames = []
    for i, f in enumerate(frames):
        assert f.dtype == np.uint8 and f.ndim == 3 and f.shape[2] == C
        t = torch.from_numpy(np.ascontiguousarray(f)).to(device, non_blocking=True)
        dev_frames.append(t)
        H, W, _ = f.shape
        _lib.call("vd3d_preprocess_describe", descs[i].ctypes.data_as(ctypes.c_void_p), t.data_ptr(), H, W, C, W * C, int(crop_top), int(size[0]), int(size[1]))
    d = torch.from_numpy(descs).to(device)
    out = torch.empty(B, C, size[0], size[1], dtype=torch.float32, device=device)
    m, s = _f32(mean), _f32(std)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.call("vd3d_preprocess", d.data_ptr(), B, C, int(size[0]), int(size[1]), vp(m), vp(s), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    out._vd3d_keepalive = (dev_frames, d)        # the frames / descriptors must outlive the asynchronous kernel
    return out
