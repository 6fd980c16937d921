"""Multi-GPU plumbing for the inference path: the per-image batch is sharded across ranks (one process per GPU,
weights and anchor tables replicated) and the ONLY exchange is one NCCL all-gather of fixed-capacity detection
records (the reference evaluates on a single GPU, R/docs/stereo3d.md:27, so parity target
is "concatenation of the per-image single-GPU results in global batch order").

Record block per rank: float32 [B_local, 1 + kmax*13]; slot 0 holds the detection count (as a float, exact below 2^24),
then kmax rows of (11 box floats, score, class).  ~53 KB per rank at B_local = 8, kmax = 128: latency-bound, so it is
a plain ncclAllGather on the compute stream right after NMS.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

REC = 13
REC_GEO = 20        # + (x3d, y3d, theta, 4 floats of the 2-D box in original-frame pixels): the post-forward columns (vd3d_post_forward)


def pack_records(results, kmax: int, device) -> torch.Tensor:
    """results: list of (scores[K], boxes[K,11], cls[K]) -> [B, 1 + kmax*13] float32 on `device`."""
    B = len(results)
    buf = torch.zeros(B, 1 + kmax * REC, dtype=torch.float32, device=device)
    for b, (s, bx, c) in enumerate(results):
        k = int(s.shape[0])
        if k > kmax:
            raise RuntimeError(f"image {b}: {k} detections exceed the all-gather record capacity {kmax}")
        buf[b, 0] = float(k)
        if k:
            rows = buf[b, 1:1 + k * REC].view(k, REC)
            rows[:, :11] = bx
            rows[:, 11] = s
            rows[:, 12] = c.to(torch.float32)
    return buf


def unpack_records(buf: torch.Tensor, rec: int = REC, geometry: Optional[list] = None) -> List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
    """[B, 1 + kmax*rec] host block -> per-image (scores, boxes[K,11], cls).  With rec == REC_GEO and a list passed as `geometry`, that
    list receives per image (box3d[K,7] = x, y, z, w, h, l, alpha in the camera frame, theta[K], box2d[K,4] in original-frame pixels)."""
    out = []
    counts = buf[:, 0].round().to(torch.int64).tolist()
    if any(k == -2 for k in counts):
        from .engine import RANGE_MSG
        from ._lib import Vd3dError
        raise Vd3dError(RANGE_MSG)
    if any(k < 0 for k in counts):
        raise RuntimeError("detection record block overflowed its capacity on some rank")
    for b, k in enumerate(counts):
        rows = buf[b, 1:1 + k * rec].view(k, rec)
        out.append((rows[:, 11].clone(), rows[:, :11].clone(), rows[:, 12].round().to(torch.int64)))
        if geometry is not None and rec >= REC_GEO:
            box3d = torch.cat([rows[:, 13:15], rows[:, 6:11]], dim=1)
            geometry.append((box3d, rows[:, 15].clone(), rows[:, 16:20].clone()))
    return out


def pack_records_device(dec, kmax: int, geometry: bool = False, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Record block built by a kernel from the fixed-capacity decode outputs (`engine.DecodeNms`): no host synchronisation.
    geometry=True appends the post-forward columns (`dec.post_forward(...)` must have run)."""
    from ._lib import call
    B = dec.B
    R = REC_GEO if geometry else REC
    rec = out if out is not None else torch.empty(B, 1 + kmax * R, dtype=torch.float32, device=dec.scores.device)
    assert rec.shape == (B, 1 + kmax * R) and rec.is_contiguous()
    if geometry:
        call("vd3d_pack_records_geo", dec.scores.data_ptr(), dec.boxes.data_ptr(), dec.cls.data_ptr(), dec.count.data_ptr(),
             dec.box3d.data_ptr(), dec.theta.data_ptr(), dec.box2d.data_ptr(), B, dec.cap, kmax, rec.data_ptr(),
             torch.cuda.current_stream().cuda_stream)
    else:
        call("vd3d_pack_records", dec.scores.data_ptr(), dec.boxes.data_ptr(), dec.cls.data_ptr(), dec.count.data_ptr(), B, dec.cap, kmax,
             rec.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return rec


def all_gather_records(rec: torch.Tensor, group=None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ONE collective: every rank's [B_local, 1 + kmax*13] block -> [world * B_local, ...] in rank order (stream-ordered NCCL)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rec
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty(world * rec.shape[0], rec.shape[1], dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec, group=group)
    return out


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard of the global batch owned by `rank` (rank r gets pairs [r*n/world, (r+1)*n/world))."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def all_gather_detections(local_results, kmax: int, device, group=None):
    """One all-gather of the padded record blocks; returns the per-image results of the GLOBAL batch on every rank.
    All ranks must hold the same number of local images."""
    buf = pack_records(local_results, kmax, device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return unpack_records(buf)
    world = dist.get_world_size(group)
    gathered = torch.empty(world * buf.shape[0], buf.shape[1], dtype=buf.dtype, device=device)
    dist.all_gather_into_tensor(gathered, buf, group=group)
    return unpack_records(gathered)
