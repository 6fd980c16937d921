"""Host-fed inference pipeline: pinned host batches -> H2D on a copy stream (double-buffered) -> forward + NMS on the compute
stream -> ONE all-gather of detection records -> asynchronous D2H into pinned memory.

The copy of batch i+1 overlaps the forward of batch i, and the host only blocks on the records of batch i-1, so the end-to-end
rate approaches the device-resident rate while every batch still pays its own H2D and D2H inside the timed region.
This is the call a data-loader loop makes (the reference's `scripts/eval.py` loop feeds one frame at a time and synchronises on
every `.cpu()`, R/networks/pipelines/testers.py).
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import parallel


class StreamedInference:
    def __init__(self, detector, batch: int, height: int, width: int, kmax: int = 512, world: int = 1, depth: int = 2):
        self.det, self.B, self.kmax, self.world, self.depth = detector, batch, kmax, world, depth
        dev = next(detector.parameters()).device
        self.dev = dev
        self.copy_stream = torch.cuda.Stream(device=dev)
        mk = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
        self.bufs = [(mk(batch, 3, height, width), mk(batch, 3, height, width), mk(batch, 3, 4)) for _ in range(depth)]
        self.host_rec = [torch.empty(world * batch, 1 + kmax * parallel.REC, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.ev_copied = [torch.cuda.Event() for _ in range(depth)]
        self.ev_free = [torch.cuda.Event() for _ in range(depth)]
        self.ev_done = [torch.cuda.Event() for _ in range(depth)]
        self.submitted = 0
        self.post_opt = bool(getattr(detector, "post_optimization", False))
        if self.post_opt and world > 1:
            raise NotImplementedError("post_optimization with the multi-GPU record all-gather needs the device kernel (vd3d_post_opt) on every rank: round 2")
        self.host_P2 = [None] * depth
        self._uncollected = [False] * depth        # slot holds a batch whose records have not been collected yet
        self.h2d_bytes = 4 * (2 * batch * 3 * height * width + batch * 12)
        self.d2h_bytes = 4 * world * batch * (1 + kmax * parallel.REC)

    def submit(self, left: torch.Tensor, right: torch.Tensor, P2: torch.Tensor) -> int:
        """Enqueue one batch given as pinned HOST tensors; returns its ticket.  Never blocks on the GPU unless `depth` batches
        are already in flight and uncollected."""
        for t in (left, right, P2):
            if t.is_cuda or not t.is_pinned():
                raise ValueError("StreamedInference.submit expects pinned host tensors")
        i = self.submitted
        k = i % self.depth
        if self._uncollected[k]:
            raise RuntimeError(f"StreamedInference: batch {i - self.depth} has not been collected; at most {self.depth} batches may be in flight")
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy_stream):
            if i >= self.depth:
                self.copy_stream.wait_event(self.ev_free[k])       # the forward that last read this staging buffer has consumed it
            dl, dr, dp = self.bufs[k]
            dl.copy_(left, non_blocking=True)
            dr.copy_(right, non_blocking=True)
            dp.copy_(P2, non_blocking=True)
            self.ev_copied[k].record(self.copy_stream)
        cur.wait_event(self.ev_copied[k])
        with torch.no_grad():
            dec = self.det.launch(dl, dr, dp)
        self.ev_free[k].record(cur)
        rec = parallel.all_gather_records(parallel.pack_records_device(dec, self.kmax))     # the single collective of the path
        self.host_rec[k].copy_(rec, non_blocking=True)
        self.ev_done[k].record(cur)
        self.host_P2[k] = P2
        self._uncollected[k] = True
        self.submitted += 1
        return i

    def collect(self, ticket: int) -> List:
        """Block until batch `ticket` is on the host; returns the per-image (scores, boxes, classes) of the GLOBAL batch."""
        k = ticket % self.depth
        if not (self.submitted - self.depth <= ticket < self.submitted) or not self._uncollected[k]:
            raise RuntimeError(f"StreamedInference: ticket {ticket} is not in flight")
        self.ev_done[k].synchronize()
        self._uncollected[k] = False
        res = parallel.unpack_records(self.host_rec[k])
        if self.post_opt:            # yaw refinement of the kept rows on the host (detectors.base.Anchor3DDetector.results does the same)
            from . import postopt
            P2h = self.host_P2[k].numpy()
            res = [(s, postopt.post_process(b, c, P2h[i]), c) if len(s) else (s, b, c) for i, (s, b, c) in enumerate(res)]
        return res
