"""Shared machinery of the anchor-based 3-D detectors (Stereo3D, Yolo3D, GroundAwareYolo3D): config parsing, lazily folded /
packed weights ("plan"), device anchor tables, the batched decode + NMS stage and the reference's list protocol."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import engine as E
from .._lib import Vd3dError
from ..anchors import AnchorTable, load_priors


class Anchor3DDetector(nn.Module):
    """Subclasses build `self.core` / `self.bbox_head` (parameter holders) and implement `build_plan(dev)` and
    `network(...) -> (cls Act, reg Act)`; everything from the head outputs to the detection triples lives here."""

    def __init__(self, network_cfg):
        super().__init__()
        self.obj_types = network_cfg["obj_types"]
        head = network_cfg["head"]
        acfg = head["anchors_cfg"]
        self.anchors_cfg = {k: acfg[k] for k in ("pyramid_levels", "strides", "sizes", "ratios", "scales")}
        # the remaining arguments of the reference's Anchors (R/heads/anchors.py:11-14) and of the head (detection_3d_head.py:30): honoured
        # or refused, never silently ignored
        y_mm = acfg.get("filter_y_threshold_min_max", (-0.5, 1.8))
        x_thr = acfg.get("filter_x_threshold", 40.0)
        if y_mm is None or x_thr is None:
            raise ValueError("anchors_cfg.filter_y_threshold_min_max / filter_x_threshold must be numbers (set test_cfg.filter_anchor=False to disable the filter)")
        self.anchor_filter = (float(y_mm[0]), float(y_mm[1]), float(x_thr))
        if int(acfg.get("anchor_prior_channel", 6)) != 6:
            raise ValueError("anchors_cfg.anchor_prior_channel must be 6 (z, sin2a, cos2a, w, h, l: the decode layout of detection_3d_head.py:218-263)")
        if not bool(head.get("read_precompute_anchor", True)):
            raise ValueError("head.read_precompute_anchor=False is not supported: the 3-D decode needs the precomputed anchor priors")
        self.num_anchors = len(acfg["pyramid_levels"]) * len(acfg["ratios"]) * len(acfg["scales"])
        self.num_classes = head["num_classes"]
        self.test_cfg = dict(head.get("test_cfg", {}))
        self.filter_anchor = bool(self.test_cfg.get("filter_anchor", head.get("loss_cfg", {}).get("filter_anchor", True)))
        lc = dict(head["layer_cfg"])
        lc.setdefault("num_anchors", self.num_anchors)
        self.layer_cfg = lc
        self.num_cls_output, self.num_reg_output = lc["num_cls_output"], lc["num_reg_output"]
        if self.num_reg_output != 12:
            raise ValueError("num_reg_output must be 12 (decode layout, detection_3d_head.py:218-263)")
        self.head_kwargs = dict(loss_cfg=dict(head.get("loss_cfg", {})),
                                num_regression_loss_terms=head.get("num_regression_loss_terms", 12), **lc)
        self.network_cfg = network_cfg
        n_rows = len(acfg["scales"]) * len(acfg["pyramid_levels"])
        self.prior_mean, self.prior_std = load_priors(head["preprocessed_path"], acfg.get("obj_types", self.obj_types),
                                                      n_rows, len(acfg["ratios"]))
        # R/heads/detection_3d_head.py:294-308 (hill climbing on the yaw): device kernel after the NMS (`decode`)
        self.post_optimization = bool(self.test_cfg.get("post_optimization", False))
        self.max_detections = int(self.test_cfg.get("max_candidates", 2048))   # fixed capacity of the decode / NMS stage
        self._plan = None
        self._plan_version = None
        self._arena = E.Arena()
        self._anchor_tables = {}
        self._decoders = {}
        self._last_decoder = None
        self.stage_hook = None            # tests: callable(name, Act-or-tensor)
        self.profile_events = None        # bench: list collecting (start, end) CUDA events of a profiled kernel

    # ---- plan (folded / packed weights) -------------------------------------------------------------------
    def _param_version(self):
        return tuple(p._version for p in self.parameters()) + tuple(b._version for b in self.buffers())

    def _device(self):
        return next(self.parameters()).device

    def build_plan(self, dev) -> dict:  # pragma: no cover
        raise NotImplementedError

    def prepare(self, force: bool = False):
        """Fold BN into conv weights, pack for the kernels, upload.  Re-run automatically when parameters change."""
        dev = self._device()
        if dev.type != "cuda":
            raise Vd3dError(f"{type(self).__name__} (B200) has no CPU path: move the module to a CUDA device first")
        ver = (self._param_version(), str(dev))
        if self._plan is not None and not force and ver == self._plan_version:
            return self._plan
        self._plan, self._plan_version = self.build_plan(dev), ver
        return self._plan

    def _hook(self, name, value):
        if self.stage_hook is not None:
            self.stage_hook(name, value)

    def _timed(self, name: str):
        """Context manager: CUDA events around a kernel IN SITU when `profile_events` is a list (bench.py sets it for the timed region; the
        events are recorded on the current stream, nothing synchronises); a no-op otherwise."""
        det = self

        class _T:
            def __enter__(self_t):
                if det.profile_events is not None:
                    self_t.e0 = torch.cuda.Event(enable_timing=True)
                    self_t.e0.record()

            def __exit__(self_t, *a):
                if det.profile_events is not None:
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record()
                    det.profile_events.append((name, self_t.e0, e1))
                return False
        return _T()

    def _anchor_table(self, H, W, dev) -> AnchorTable:
        key = (H, W, str(dev))
        if key not in self._anchor_tables:
            self._anchor_tables[key] = AnchorTable((H, W), self.anchors_cfg, self.prior_mean, self.prior_std, dev)
        return self._anchor_tables[key]

    # ---- head outputs -> detections -------------------------------------------------------------------------
    def decode(self, cls: E.Act, reg: E.Act, P2: torch.Tensor, H: int, W: int) -> E.DecodeNms:
        """get_anchor + get_bboxes (R/heads/detection_3d_head.py:310-321,341-400), batched, no host sync."""
        dev = cls.t.device
        B = cls.B
        tab = self._anchor_table(H, W, dev)
        N = tab.N
        assert cls.H * cls.W * cls.C == N * self.num_cls_output and reg.C * reg.H * reg.W == N * 12
        assert cls.co == 0 and reg.co == 0 and cls.cs == cls.C and reg.cs == reg.C
        mask = self._arena.get("mask", (B, N), dev, dtype=torch.uint8)
        if self.filter_anchor:
            E.anchor_mask(tab.anchors, tab.means_z, P2, mask, *self.anchor_filter)
        else:
            mask.fill_(1)
        self._hook("mask", mask)
        key = (B, str(dev))
        if key not in self._decoders:
            self._decoders[key] = E.DecodeNms(B, self.max_detections, dev)
        dec = self._decoders[key]
        dec.run(cls.t.view(B, N, self.num_cls_output), reg.t.view(B, N, 12), tab.anchors, tab.mean_std, mask,
                self.num_classes, self.test_cfg.get("score_thr", 0.5), self.test_cfg.get("nms_iou_thr", 0.5), W, H)
        if self.post_optimization:
            # head.test_cfg.post_optimization (R/heads/detection_3d_head.py:294-308): yaw hill climbing of the kept car boxes deeper than
            # 3 m, in place on the fixed-capacity NMS output, stream-ordered (the reference searches on the CPU with one `.item()` per box)
            dec.post_opt(P2)
        self._last_decoder = dec
        return dec

    @staticmethod
    def results(dec: E.DecodeNms):
        """Per-image (scores, bboxes, cls) triples (one D2H read of the counts)."""
        return [(s.clone(), b.clone(), c.clone()) for (s, b, c) in dec.results()]

    def train_forward(self, *a, **k):
        raise NotImplementedError("training forward is out of scope of the B200 inference path (SURVEY.md section 2)")


def synth_load(det: nn.Module, seed: int):
    """Fill a detector with the seeded synthetic weights (visualdet3d_b200.synth); returns the state_dict used."""
    from .. import synth
    shapes = {k: tuple(v.shape) for k, v in det.state_dict().items()}
    sd = synth.synth_state_dict(shapes, seed, cls_gain=synth.CLS_GAIN.get(type(det).__name__, 1.6))
    det.load_state_dict(sd, strict=False)
    return sd
