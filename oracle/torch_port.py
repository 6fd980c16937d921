"""ORACLE (test infrastructure, NOT product code).

A CPU restatement, in plain fp32 PyTorch functional ops, of the visualDet3D inference forward named by
BASELINE.json's north_star.  Every function cites the reference file:line it follows
(R/ = /root/reference/visualDet3D/networks).  It operates on a reference-format ``state_dict`` so the same
weights drive the reference, this port and the CUDA path.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified reference in the build
container, runs it on seeded inputs/weights and commits the outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this port against those fixtures.  Arithmetic that lives in third-party
code (oneDNN convs, torchvision.ops.nms, F.grid_sample) is called through the same torch/torchvision entry
points the reference calls, so "same inputs -> same outputs" holds on one machine.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this file.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------------------
def bn(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """nn.BatchNorm{2,3}d in eval mode (running stats), eps = 1e-5 (torch default; the reference never changes it)."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=1e-5)


def conv(sd: SD, p: str, x: torch.Tensor, stride=1, padding=0, dilation=1, groups=1) -> torch.Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding,
                    dilation=dilation, groups=groups)


def basic_block(sd: SD, p: str, x: torch.Tensor, stride: int = 1, dilation: int = 1) -> torch.Tensor:
    """R/backbones/resnet.py:23-52 (BasicBlock.forward)."""
    out = F.relu(bn(sd, p + ".bn1", conv(sd, p + ".conv1", x, stride=stride, padding=1)))
    out = bn(sd, p + ".bn2", conv(sd, p + ".conv2", out, padding=dilation, dilation=dilation))
    if (p + ".downsample.0.weight") in sd:
        x = bn(sd, p + ".downsample.1", conv(sd, p + ".downsample.0", x, stride=stride))
    return F.relu(out + x)


def bottleneck(sd: SD, p: str, x: torch.Tensor, stride: int = 1, dilation: int = 1) -> torch.Tensor:
    """R/backbones/resnet.py:55-91 (Bottleneck.forward)."""
    out = F.relu(bn(sd, p + ".bn1", conv(sd, p + ".conv1", x)))
    out = F.relu(bn(sd, p + ".bn2", conv(sd, p + ".conv2", out, stride=stride, padding=dilation, dilation=dilation)))
    out = bn(sd, p + ".bn3", conv(sd, p + ".conv3", out))
    if (p + ".downsample.0.weight") in sd:
        x = bn(sd, p + ".downsample.1", conv(sd, p + ".downsample.0", x, stride=stride))
    return F.relu(out + x)


RESNET_LAYERS = {18: ("basic", [2, 2, 2, 2]), 34: ("basic", [3, 4, 6, 3]), 50: ("bottle", [3, 4, 6, 3]),
                 101: ("bottle", [3, 4, 23, 3]), 152: ("bottle", [3, 8, 36, 3])}


def resnet(sd: SD, p: str, img: torch.Tensor, depth: int, num_stages: int = 4, out_indices=(-1, 0, 1, 2, 3),
           strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1)) -> List[torch.Tensor]:
    """R/backbones/resnet.py:184-198 (ResNet.forward); stage layout `_make_layer` :140-152.
    Note `_make_layer` gives the FIRST block of a stage the stride and dilation=1, later blocks dilation."""
    kind, layers = RESNET_LAYERS[depth]
    block = basic_block if kind == "basic" else bottleneck
    outs = []
    x = F.relu(bn(sd, p + ".bn1", conv(sd, p + ".conv1", img, stride=2, padding=3)))
    if -1 in out_indices:
        outs.append(x)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for i in range(num_stages):
        for j in range(layers[i]):
            x = block(sd, f"{p}.layer{i + 1}.{j}", x, stride=strides[i] if j == 0 else 1,
                      dilation=1 if j == 0 else dilations[i])
        if i in out_indices:
            outs.append(x)
    return outs


def res_ghost(sd: SD, p: str, x: torch.Tensor, oup: int) -> torch.Tensor:
    """R/lib/ghost_module.py:46-64 (ResGhostModule.forward, stride 1): out = cat[x, x1, x2][:, :oup]."""
    x1 = F.relu(bn(sd, p + ".primary_conv.2", conv(sd, p + ".primary_conv.1", x, padding=sd[p + ".primary_conv.1.weight"].shape[-1] // 2)))
    w2 = sd[p + ".cheap_operation.0.weight"]
    x2 = F.relu(bn(sd, p + ".cheap_operation.1",
                   F.conv2d(x1, w2, None, stride=1, padding=w2.shape[-1] // 2, groups=x1.shape[1])))
    return torch.cat([x, x1, x2], dim=1)[:, :oup]


# --------------------------------------------------------------------------------------------------------
# cost volumes  (R/lib/PSM_cost_volume.py)
# --------------------------------------------------------------------------------------------------------
def psm_cosine(left: torch.Tensor, right: torch.Tensor, max_disp: int, downsample_scale: int) -> torch.Tensor:
    """R/lib/PSM_cost_volume.py:76-91: cost[b,i,h,w] = mean_c L[b,c,h,w] * R[b,c,h,w-i] for w >= i else 0."""
    D = int(max_disp / downsample_scale)
    B, C, H, W = left.shape
    cost = torch.zeros(B, D, H, W, dtype=left.dtype)
    for i in range(D):
        if i > 0:
            if i < W:
                cost[:, i, :, i:] = (left[:, :, :, i:] * right[:, :, :, :-i]).mean(dim=1)
        else:
            cost[:, i] = (left * right).mean(dim=1)
    return cost


def concat_volume(lf: torch.Tensor, rf: torch.Tensor, D: int) -> torch.Tensor:
    """R/lib/PSM_cost_volume.py:44-60: [B, 2F, D, H, W]; plane i holds L' (w>=i) and R' shifted by i."""
    B, Fc, H, W = lf.shape
    cost = torch.zeros(B, 2 * Fc, D, H, W, dtype=lf.dtype)
    for i in range(D):
        if i > 0:
            if i < W:
                cost[:, :Fc, i, :, i:] = lf[:, :, :, i:]
                cost[:, Fc:, i, :, i:] = rf[:, :, :, :-i]
        else:
            cost[:, :Fc, i] = lf
            cost[:, Fc:, i] = rf
    return cost


def cost_volume(sd: SD, p: str, left: torch.Tensor, right: torch.Tensor, max_disp: int, downsample_scale: int) -> torch.Tensor:
    """R/lib/PSM_cost_volume.py:40-63 (CostVolume.forward)."""
    D = int(max_disp / downsample_scale)
    B, _, H, W = left.shape
    lf = F.relu(bn(sd, p + ".down_sample.1", conv(sd, p + ".down_sample.0", left)))
    rf = F.relu(bn(sd, p + ".down_sample.1", conv(sd, p + ".down_sample.0", right)))
    cost = concat_volume(lf, rf, D)
    cost = F.relu(bn(sd, p + ".conv3d.1", F.conv3d(cost, sd[p + ".conv3d.0.weight"], sd[p + ".conv3d.0.bias"], padding=1)))
    cost = F.relu(bn(sd, p + ".conv3d.4", F.conv3d(cost, sd[p + ".conv3d.3.weight"], sd[p + ".conv3d.3.bias"], padding=1)))
    return cost.reshape(B, -1, H, W).contiguous()


def cost_volume_pyramid(sd: SD, p: str, v4: torch.Tensor, v8: torch.Tensor, v16: torch.Tensor) -> torch.Tensor:
    """R/detectors/yolostereo3d_core.py:63-71 (CostVolumePyramid.forward, eval branch)."""
    c4 = v4.shape[1]
    x = res_ghost(sd, p + ".four_to_eight.0", v4, 3 * c4)
    x = F.avg_pool2d(x, 2)
    x = basic_block(sd, p + ".four_to_eight.2", x)
    v8 = torch.cat([x, v8], dim=1)
    x = res_ghost(sd, p + ".eight_to_sixteen.0", v8, 3 * v8.shape[1])
    x = F.avg_pool2d(x, 2)
    x = basic_block(sd, p + ".eight_to_sixteen.2", x)
    v16 = torch.cat([x, v16], dim=1)
    x = res_ghost(sd, p + ".depth_reason.0", v16, 3 * v16.shape[1])
    x = basic_block(sd, p + ".depth_reason.1", x)
    return x


def stereo_core(sd: SD, left: torch.Tensor, right: torch.Tensor, depth: int = 34, stages: dict | None = None) -> torch.Tensor:
    """R/detectors/yolostereo3d_core.py:110-126 + StereoMerging.forward :88-94. Returns features [B,1408,H/16,W/16]."""
    B = left.shape[0]
    feats = resnet(sd, "core.backbone", torch.cat([left, right], dim=0), depth, num_stages=3, out_indices=(0, 1, 2))
    lf = [f[:B] for f in feats]
    rf = [f[B:] for f in feats]
    v0 = psm_cosine(lf[0], rf[0], 96, 4)
    v1 = psm_cosine(lf[1], rf[1], 192, 8)
    v2 = cost_volume(sd, "core.neck.cost_volume_2", lf[2], rf[2], 192, 16)
    psv = cost_volume_pyramid(sd, "core.neck.depth_reasoning", v0, v1, v2)
    features = torch.cat([lf[2], psv], dim=1)
    if stages is not None:
        stages.update(feat4=feats[0], feat8=feats[1], feat16=feats[2], vol4=v0, vol8=v1, vol16=v2, psv=psv, features=features)
    return features


def anchor_flatten(x: torch.Tensor, c: int) -> torch.Tensor:
    """R/lib/blocks.py:117-136."""
    return x.permute(0, 2, 3, 1).contiguous().view(x.shape[0], -1, c)


def stereo_head(sd: SD, features: torch.Tensor, num_cls_output: int, num_reg_output: int = 12) -> Tuple[torch.Tensor, torch.Tensor]:
    """R/heads/detection_3d_head.py:500-533 (StereoHead.init_layers) + :84-88 (forward); Dropout2d is identity in eval."""
    p = "bbox_head.cls_feature_extraction"
    x = F.relu(conv(sd, p + ".0", features, padding=1))
    x = F.relu(conv(sd, p + ".3", x, padding=1))
    cls = anchor_flatten(conv(sd, p + ".6", x, padding=1), num_cls_output)
    p = "bbox_head.reg_feature_extraction"
    x = F.relu(bn(sd, p + ".0.sequence.1", conv(sd, p + ".0.sequence.0", features, padding=1)))  # ConvBnReLU (blocks.py:24-43)
    x = F.relu(basic_block(sd, p + ".1", x))
    reg = anchor_flatten(conv(sd, p + ".3", x, padding=1), num_reg_output)
    return cls, reg


# --------------------------------------------------------------------------------------------------------
# anchors  (R/heads/anchors.py)
# --------------------------------------------------------------------------------------------------------
def generate_anchors(base_size: float, ratios: np.ndarray, scales: np.ndarray) -> np.ndarray:
    """R/heads/anchors.py:152-183: float64, per-cell order a = ratio_idx * n_scales + scale_idx."""
    ratios = np.asarray(ratios, dtype=np.float64)
    scales = np.asarray(scales, dtype=np.float64)
    rr = np.repeat(ratios, len(scales))
    ss = np.tile(scales, len(ratios))
    side = base_size * ss
    areas = side * side
    w = np.sqrt(areas / rr)
    h = w * rr
    out = np.zeros((len(rr), 4))
    out[:, 0] = 0 - w * 0.5
    out[:, 1] = 0 - h * 0.5
    out[:, 2] = w - w * 0.5
    out[:, 3] = h - h * 0.5
    return out


def shift_anchors(shape_hw: Sequence[int], stride: float, anchors: np.ndarray) -> np.ndarray:
    """R/heads/anchors.py:219-239: n = (y*W + x)*A + a."""
    sx = (np.arange(0, shape_hw[1]) + 0.5) * stride
    sy = (np.arange(0, shape_hw[0]) + 0.5) * stride
    sx, sy = np.meshgrid(sx, sy)
    shifts = np.stack([sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel()], axis=1)
    return (anchors[None, :, :] + shifts[:, None, :]).reshape(-1, 4)


def build_anchors(image_hw: Sequence[int], acfg: dict, prior_mean: np.ndarray, prior_std: np.ndarray):
    """R/heads/anchors.py:59-91. prior_* : [types, n_scales*levels, n_ratios, 6] float64.
    Returns anchors f32 [N,4], anchor_mean_std f32 [N, types, 6, 2], and float64 -> f32 prior means [types, N, 6]."""
    levels, strides, sizes = acfg["pyramid_levels"], acfg["strides"], acfg["sizes"]
    ratios, scales = np.asarray(acfg["ratios"], dtype=np.float64), np.asarray(acfg["scales"], dtype=np.float64)
    image_shape = np.array(image_hw)
    all_anchors = np.zeros((0, 4)).astype(np.float32)
    for idx, lv in enumerate(levels):
        shp = (image_shape + 2 ** lv - 1) // (2 ** lv)
        all_anchors = np.append(all_anchors, shift_anchors(shp, strides[idx], generate_anchors(sizes[idx], ratios, scales)), axis=0)
    # anchors2indexes :45-57 (float64)
    sz = np.sqrt((all_anchors[:, 2] - all_anchors[:, 0]) * (all_anchors[:, 3] - all_anchors[:, 1]))
    sizes_int = np.argmin(np.abs(sz - (np.array(sizes) * scales)[:, None]), axis=0)
    rt = (all_anchors[:, 3] - all_anchors[:, 1]) / (all_anchors[:, 2] - all_anchors[:, 0])
    ratio_int = np.argmin(np.abs(rt - ratios[:, None]), axis=0)
    means = torch.tensor(prior_mean[:, sizes_int, ratio_int], dtype=torch.float32)  # [types, N, 6]   (image.new(float64 ndarray) casts to f32)
    stds = torch.tensor(prior_std[:, sizes_int, ratio_int], dtype=torch.float32)
    mean_std = torch.stack([means, stds], dim=-1).permute(1, 0, 2, 3).contiguous()  # [N, types, 6, 2]
    anchors = torch.tensor(all_anchors.astype(np.float32))
    return anchors, mean_std, means


def useful_mask(anchors: torch.Tensor, means: torch.Tensor, P2: torch.Tensor,
                y_min_max=(-0.5, 1.8), x_thr=40.0) -> torch.Tensor:
    """R/heads/anchors.py:93-111. anchors [N,4] f32, means [types,N,6] f32, P2 [B,3,4] -> bool [B,N]."""
    xc = anchors[:, 0:4:2].mean(dim=1)
    yc = anchors[:, 1:4:2].mean(dim=1)
    fy = P2[:, 1:2, 1:2]
    cy = P2[:, 1:2, 2:3]
    cx = P2[:, 0:1, 2:3]
    z = means[:, :, 0]
    x3d = (xc * z - cx * z) / fy
    y3d = (yc * z - cy * z) / fy
    return torch.any((y3d > y_min_max[0]) * (y3d < y_min_max[1]) * (x3d.abs() < x_thr), dim=1)


# --------------------------------------------------------------------------------------------------------
# decode + NMS  (R/heads/detection_3d_head.py:218-263, 341-400)
# --------------------------------------------------------------------------------------------------------
def decode(boxes, deltas, mean_std, label, alpha_score):
    """R/heads/detection_3d_head.py:218-263 (_decode)."""
    std = torch.tensor([0.1, 0.1, 0.2, 0.2, 0.1, 0.1, 1, 1, 1, 1, 1, 1], dtype=torch.float32)
    widths = boxes[..., 2] - boxes[..., 0]
    heights = boxes[..., 3] - boxes[..., 1]
    ctr_x = boxes[..., 0] + 0.5 * widths
    ctr_y = boxes[..., 1] + 0.5 * heights
    dx, dy = deltas[..., 0] * std[0], deltas[..., 1] * std[1]
    dw, dh = deltas[..., 2] * std[2], deltas[..., 3] * std[3]
    pcx = ctr_x + dx * widths
    pcy = ctr_y + dy * heights
    pw = torch.exp(dw) * widths
    ph = torch.exp(dh) * heights
    x1, y1, x2, y2 = pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph
    sel = mean_std[torch.arange(len(label)), label]  # == anchors_3d_mean_std[one_hot_mask]
    mask = sel[:, 0, 0] > 0
    cx1 = ctr_x + deltas[..., 4] * std[4] * widths
    cy1 = ctr_y + deltas[..., 5] * std[5] * heights
    z = deltas[..., 6] * sel[:, 0, 1] + sel[:, 0, 0]
    sn = deltas[..., 7] * sel[:, 1, 1] + sel[:, 1, 0]
    cs = deltas[..., 8] * sel[:, 2, 1] + sel[:, 2, 0]
    alpha = torch.atan2(sn, cs) / 2.0
    w3 = deltas[..., 9] * sel[:, 3, 1] + sel[:, 3, 0]
    h3 = deltas[..., 10] * sel[:, 4, 1] + sel[:, 4, 0]
    l3 = deltas[..., 11] * sel[:, 5, 1] + sel[:, 5, 0]
    out = torch.stack([x1, y1, x2, y2, cx1, cy1, z, w3, h3, l3, alpha], dim=1)
    out[alpha_score[:, 0] < 0.5, -1] += np.pi
    return out, mask


def get_bboxes(cls_preds, reg_preds, anchors, mean_std, mask, image_hw, num_classes, score_thr, nms_iou_thr,
               stages: dict | None = None):
    """R/heads/detection_3d_head.py:341-400 for ONE image (cls_preds [N,C+1], reg_preds [N,12], mask [N] bool).
    Class-agnostic NMS always (typo `cls_agnositc` at :381)."""
    from torchvision.ops import nms
    p = cls_preds.sigmoid()
    cls_score = p[..., 0:num_classes][mask]
    alpha_score = p[..., num_classes:num_classes + 1][mask]
    reg = reg_preds[mask]
    anc = anchors[mask]
    ms = mean_std[mask]
    idx0 = torch.nonzero(mask)[:, 0]
    max_score, label = cls_score.max(dim=-1)
    hs = max_score > score_thr
    anc, ms, alpha_score, reg, max_score, label, idx0 = anc[hs], ms[hs], alpha_score[hs], reg[hs], max_score[hs], label[hs], idx0[hs]
    boxes, valid = decode(anc, reg, ms, label, alpha_score)
    H, W = image_hw
    boxes[:, 0] = torch.clamp(boxes[:, 0], min=0)   # R/utils/utils.py:181-196 ClipBoxes
    boxes[:, 1] = torch.clamp(boxes[:, 1], min=0)
    boxes[:, 2] = torch.clamp(boxes[:, 2], max=W)
    boxes[:, 3] = torch.clamp(boxes[:, 3], max=H)
    max_score, boxes, label, idx0 = max_score[valid], boxes[valid], label[valid], idx0[valid]
    keep = nms(boxes[:, :4], max_score, nms_iou_thr)
    if stages is not None:
        stages.update(cand_anchor_idx=idx0.clone(), cand_scores=max_score.clone(), cand_boxes=boxes.clone(),
                      cand_labels=label.clone(), keep=keep.clone())
    return max_score[keep], boxes[keep], label[keep], idx0[keep]


# --------------------------------------------------------------------------------------------------------
# detectors
# --------------------------------------------------------------------------------------------------------
def stereo3d_forward(sd: SD, left, right, P2, cfg: dict, prior_mean, prior_std, stages: dict | None = None):
    """R/detectors/yolostereo3d_detector.py:77-96 (Stereo3D.test_forward), looped per image for B > 1
    (the reference asserts B == 1; core + head are batch-invariant in eval mode)."""
    with torch.no_grad():
        feats = stereo_core(sd, left, right, cfg["backbone"]["depth"], stages)
        ncls = cfg["head"]["num_classes"]
        cls_preds, reg_preds = stereo_head(sd, feats, ncls + 1)
        anchors, mean_std, means = build_anchors(left.shape[2:], cfg["head"]["anchors_cfg"], prior_mean, prior_std)
        mask = useful_mask(anchors, means, P2)
        if stages is not None:
            stages.update(cls_preds=cls_preds, reg_preds=reg_preds, anchors=anchors, mean_std=mean_std, mask=mask)
        outs = []
        for b in range(left.shape[0]):
            st = {} if stages is not None else None
            outs.append(get_bboxes(cls_preds[b], reg_preds[b], anchors, mean_std, mask[b], left.shape[2:], ncls,
                                   cfg["head"]["test_cfg"]["score_thr"], cfg["head"]["test_cfg"]["nms_iou_thr"], st))
            if stages is not None:
                stages.setdefault("per_image", []).append(st)
        return outs


# --------------------------------------------------------------------------------------------------------
# deformable conv / iou3d restatements (CPU).  The GPU oracle for these two ops is the reference's own compiled
# extension (oracle/build_ref.py -> oracle/_ref); these CPU versions are the independent second opinion.
# --------------------------------------------------------------------------------------------------------
def modulated_deform_conv(x, offset, mask, weight, bias=None, stride=1, padding=0, dilation=1):
    """R/lib/ops/dcn/src/cuda/deform_conv_cuda_kernel.cu:467-497,570-633 == torchvision.ops.deform_conv2d (same mmcv
    lineage: (dh, dw)-interleaved offsets, `> -1 / < H` validity, zero outside) — SURVEY.md section 8(c).6."""
    import torchvision
    return torchvision.ops.deform_conv2d(x, offset, weight, bias, stride=stride, padding=padding, dilation=dilation, mask=mask)


def modulated_deform_conv_pack(sd: SD, p: str, x, stride=1, padding=1, dilation=1):
    """ModulatedDeformConvPack.forward (R/lib/ops/dcn/deform_conv.py:459-466)."""
    out = conv(sd, p + ".conv_offset", x, stride=stride, padding=padding, dilation=dilation)      # conv_offset shares the dilation (:441-449)
    o1, o2, m = torch.chunk(out, 3, dim=1)
    return modulated_deform_conv(x, torch.cat((o1, o2), dim=1), torch.sigmoid(m), sd[p + ".weight"], sd.get(p + ".bias"),
                                 stride, padding, dilation)


def rotated_overlap_bev(box_a: np.ndarray, box_b: np.ndarray) -> float:
    """Area of intersection of two rotated rectangles [x1,y1,x2,y2,ry] by Sutherland-Hodgman clipping in float64 — an
    independent algorithm from R/lib/ops/iou3d/src/iou3d_kernel.cu:108-212 (edge intersections + contained corners +
    angular sort); corner convention (rotation about the centre with (cos, sin; -sin, cos)) from :99-103,124-146."""
    def corners(b):
        x1, y1, x2, y2, a = [float(v) for v in b]
        cx, cy = (x1 + x2) / 2, (y1 + y2) / 2
        c, s = math.cos(a), math.sin(a)
        pts = []
        for (px, py) in ((x1, y1), (x2, y1), (x2, y2), (x1, y2)):
            pts.append(((px - cx) * c + (py - cy) * s + cx, -(px - cx) * s + (py - cy) * c + cy))
        return pts

    def area(poly):
        return 0.5 * sum(poly[i][0] * poly[(i + 1) % len(poly)][1] - poly[(i + 1) % len(poly)][0] * poly[i][1] for i in range(len(poly)))

    subj, clip = corners(box_a), corners(box_b)
    if area(clip) < 0:
        clip = clip[::-1]
    out = subj
    for i in range(4):
        a, b = clip[i], clip[(i + 1) % 4]
        inp, out = out, []
        if not inp:
            break
        side = lambda p: (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
        for j in range(len(inp)):
            p, q = inp[j], inp[(j + 1) % len(inp)]
            sp, sq = side(p), side(q)
            if sp >= 0:
                out.append(p)
            if (sp >= 0) != (sq >= 0):
                t = sp / (sp - sq)
                out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
    return abs(area(out)) if len(out) >= 3 else 0.0


# --------------------------------------------------------------------------------------------------------
# monocular detectors (R/detectors/yolomono3d_detector.py, R/lib/look_ground.py)
# --------------------------------------------------------------------------------------------------------
def look_ground(sd: SD, p: str, x: torch.Tensor, P2: torch.Tensor, baseline=0.54, relative_elevation=1.65) -> torch.Tensor:
    """LookGround.forward (R/lib/look_ground.py:24-71), same op order."""
    P2 = P2.clone()
    P2[:, 0:2] /= 16.0
    disp = torch.tanh(conv(sd, p + ".disp_create.0", x, padding=1))
    disp = 0.1 * (0.05 * disp + 0.95 * disp)
    B, _, H, W = x.shape
    yy = torch.arange(H, dtype=torch.float32).view(1, H, 1).expand(1, H, W)
    fy, cy, Ty = P2[:, 1:2, 1:2], P2[:, 1:2, 2:3], P2[:, 1:2, 3:4]
    disparity = F.relu(fy * baseline * (yy - cy) / (torch.abs(fy * relative_elevation + Ty) + 1e-10))
    x_base = torch.linspace(-1, 1, W).repeat(B, H, 1).type_as(x)
    y_base = torch.linspace(-1, 1, H).repeat(B, W, 1).transpose(1, 2).type_as(x)
    h_mean = 1.535
    y_shifts_base = F.relu(h_mean * (yy - cy) / (2 * (relative_elevation - 0.5 * h_mean))) / (yy.shape[1] * 0.5)
    y_shifts = y_shifts_base + disp[:, 0, :, :]
    flow = torch.stack((x_base, y_base + y_shifts), dim=3)
    feats = torch.cat([disparity.unsqueeze(1), x], dim=1)
    out = F.grid_sample(feats, flow, mode="bilinear", padding_mode="border", align_corners=True)
    return F.relu(x + conv(sd, p + ".extract", out) * sd[p + ".alpha"])


def mono_cls_tower(sd: SD, feats: torch.Tensor, ncls_out: int) -> torch.Tensor:
    p = "bbox_head.cls_feature_extraction"
    x = F.relu(conv(sd, p + ".0", feats, padding=1))
    x = F.relu(conv(sd, p + ".3", x, padding=1))
    return anchor_flatten(conv(sd, p + ".6", x, padding=1), ncls_out)


def yolo3d_head(sd: SD, feats: torch.Tensor, ncls_out: int):
    """AnchorBasedDetection3DHead (R/heads/detection_3d_head.py:47-88): DCNv2+BN+ReLU, conv+BN+ReLU, conv."""
    cls = mono_cls_tower(sd, feats, ncls_out)
    p = "bbox_head.reg_feature_extraction"
    x = F.relu(bn(sd, p + ".1", modulated_deform_conv_pack(sd, p + ".0", feats, 1, 1, 1)))
    x = F.relu(bn(sd, p + ".4", conv(sd, p + ".3", x, padding=1)))
    return cls, anchor_flatten(conv(sd, p + ".6", x, padding=1), 12)


def gac_head(sd: SD, feats: torch.Tensor, P2: torch.Tensor, ncls_out: int, stages: dict | None = None):
    """GroundAwareHead (R/detectors/yolomono3d_detector.py:12-53)."""
    cls = mono_cls_tower(sd, feats, ncls_out)
    p = "bbox_head.reg_feature_extraction"
    x = look_ground(sd, p + ".0", feats, P2)
    if stages is not None:
        stages["gac"] = x
    x = F.relu(bn(sd, p + ".2", conv(sd, p + ".1", x, padding=1)))
    x = F.relu(bn(sd, p + ".5", conv(sd, p + ".4", x, padding=1)))
    return cls, anchor_flatten(conv(sd, p + ".7", x, padding=1), 12)


def mono3d_forward(sd: SD, images, P2, cfg: dict, prior_mean, prior_std, stages: dict | None = None):
    """Yolo3D / GroundAwareYolo3D test_forward (R/detectors/yolomono3d_detector.py:100-120), looped per image for B > 1."""
    with torch.no_grad():
        bb = cfg["backbone"]
        feats = resnet(sd, "core.backbone", images, bb["depth"], num_stages=bb["num_stages"], out_indices=tuple(bb["out_indices"]))[0]
        ncls = cfg["head"]["num_classes"]
        if cfg["name"] == "GroundAwareYolo3D":
            cls_preds, reg_preds = gac_head(sd, feats, P2, ncls + 1, stages)
        else:
            cls_preds, reg_preds = yolo3d_head(sd, feats, ncls + 1)
        anchors, mean_std, means = build_anchors(images.shape[2:], cfg["head"]["anchors_cfg"], prior_mean, prior_std)
        mask = useful_mask(anchors, means, P2)
        if stages is not None:
            stages.update(features=feats, cls_preds=cls_preds, reg_preds=reg_preds, anchors=anchors, mean_std=mean_std, mask=mask)
        outs = []
        for b in range(images.shape[0]):
            st = {} if stages is not None else None
            outs.append(get_bboxes(cls_preds[b], reg_preds[b], anchors, mean_std, mask[b], images.shape[2:], ncls,
                                   cfg["head"]["test_cfg"]["score_thr"], cfg["head"]["test_cfg"]["nms_iou_thr"], st))
            if stages is not None:
                stages.setdefault("per_image", []).append(st)
        return outs


# --------------------------------------------------------------------------------------------------------
# DLA-34 + DLA up-sampling + MonoFlex head  (R/backbones/dla.py, dla_utils.py, R/heads/monoflex_head.py)
# --------------------------------------------------------------------------------------------------------
def dla_block(sd: SD, p: str, x, residual=None, stride=1):
    """BasicBlock.forward (R/backbones/dla.py:56-70)."""
    if residual is None:
        residual = x
    out = F.relu(bn(sd, p + ".bn1", conv(sd, p + ".conv1", x, stride=stride, padding=1)))
    out = bn(sd, p + ".bn2", conv(sd, p + ".conv2", out, padding=1))
    return F.relu(out + residual)


def dla_tree(sd: SD, p: str, x, levels: int, stride: int, level_root: bool, children=None):
    """Tree.forward (R/backbones/dla.py:216-230); root_residual is False for DLA-34."""
    children = [] if children is None else children
    bottom = F.max_pool2d(x, stride, stride=stride) if stride > 1 else x
    residual = bn(sd, p + ".project.1", conv(sd, p + ".project.0", bottom)) if (p + ".project.0.weight") in sd else bottom
    if level_root:
        children.append(bottom)
    if levels == 1:
        x1 = dla_block(sd, p + ".tree1", x, residual, stride)
        x2 = dla_block(sd, p + ".tree2", x1)
        k = sd[p + ".root.conv.weight"].shape[-1]
        return F.relu(bn(sd, p + ".root.bn", conv(sd, p + ".root.conv", torch.cat([x2, x1] + children, 1), padding=(k - 1) // 2)))
    x1 = dla_tree(sd, p + ".tree1", x, levels - 1, stride, False)          # (the `residual` argument is overwritten inside, :219)
    children.append(x1)
    return dla_tree(sd, p + ".tree2", x1, levels - 1, 1, False, children)


def dla34(sd: SD, p: str, img):
    """DLA.forward (R/backbones/dla.py:317-326) for dla34, out_indices (0..5)."""
    x = F.relu(bn(sd, p + ".base_layer.1", conv(sd, p + ".base_layer.0", img, padding=3)))
    ys = []
    x = F.relu(bn(sd, p + ".level0.1", conv(sd, p + ".level0.0", x, padding=1))); ys.append(x)
    x = F.relu(bn(sd, p + ".level1.1", conv(sd, p + ".level1.0", x, stride=2, padding=1))); ys.append(x)
    for i, (lv, root) in enumerate([(1, False), (2, True), (2, True), (1, True)]):
        x = dla_tree(sd, f"{p}.level{i + 2}", x, lv, 2, root)
        ys.append(x)
    return ys


def dla_deform(sd: SD, p: str, x):
    """DeformConv.forward (R/backbones/dla_utils.py:52-56)."""
    return F.relu(bn(sd, p + ".actf.0", modulated_deform_conv_pack(sd, p + ".conv", x, 1, 1, 1)))


def ida_up(sd: SD, p: str, layers, startp, endp):
    """IDAUp.forward (R/backbones/dla_utils.py:79-85)."""
    for i in range(startp + 1, endp):
        k = i - startp
        w = sd[f"{p}.up_{k}.weight"]
        f = w.shape[-1] // 2
        up = F.conv_transpose2d(dla_deform(sd, f"{p}.proj_{k}", layers[i]), w, None, stride=f, padding=f // 2, groups=w.shape[0])
        layers[i] = dla_deform(sd, f"{p}.node_{k}", up + layers[i - 1])


def dla_seg_upsample(sd: SD, p: str, tensors, first_level=2, last_level=5):
    """DLASegUpsample.forward (R/backbones/dla_utils.py:147-155) incl. DLAUp.forward (:106-112)."""
    layers = list(tensors)
    out = [layers[-1]]
    for i in range(len(layers) - first_level - 1):
        ida_up(sd, f"{p}.dla_up.ida_{i}", layers, len(layers) - i - 2, len(layers))
        out.insert(0, layers[-1])
    y = [out[i].clone() for i in range(last_level - first_level)]
    ida_up(sd, p + ".ida_up", y, 0, len(y))
    return y[-1]


def km3d_heads(sd: SD, feat, names):
    """KM3DHead.forward (R/heads/km3d_head.py:353-357)."""
    return {n: conv(sd, f"bbox_head.head_layers.{n}.2", F.relu(conv(sd, f"bbox_head.head_layers.{n}.0", feat, padding=1))) for n in names}


def monoflex_get_bboxes(output, P2, image_hw, score_thr=0.1, nms_iou_thr=0.5, unc_range=(-10, 10), K=100):
    """MonoFlexHead.get_bboxes (R/heads/monoflex_head.py:114-179) for ONE image (output maps [1, n, H, W])."""
    from torchvision.ops import nms
    hm = torch.sigmoid(output["hm"])
    hmax = F.max_pool2d(hm, (3, 3), stride=1, padding=1)
    heat = hm * (hmax == hm).float()
    batch, cat, height, width = heat.shape
    topk_scores, topk_inds = torch.topk(heat.view(batch, cat, -1), K)
    topk_inds = topk_inds % (height * width)
    topk_ys = (topk_inds / width).int().float()
    topk_xs = (topk_inds % width).int().float()
    scores, topk_ind = torch.topk(topk_scores.view(batch, -1), K)
    clses = (topk_ind / K).int()
    g = lambda t: t.view(batch, -1, 1).gather(1, topk_ind.unsqueeze(2)).view(batch, K)
    inds, ys, xs = g(topk_inds), g(topk_ys), g(topk_xs)
    gather = lambda name: output[name].permute(0, 2, 3, 1).reshape(batch, height * width, -1).gather(
        1, inds.long().unsqueeze(2).expand(batch, K, output[name].shape[1]))[0]
    scores, clses, ys, xs = scores[0], clses[0], ys[0], xs[0]
    reg2d = gather("bbox2d")
    bbox2d = torch.stack([xs - reg2d[:, 0], ys - reg2d[:, 1], xs + reg2d[:, 2], ys + reg2d[:, 3]], dim=-1)
    depth_decoded = torch.exp(-gather("depth"))
    hps = gather("hps").reshape(K, -1, 2)
    dim = gather("dim")
    ph = dim[..., 1]
    f = P2[0, 0, 0]
    center_h = hps[..., -2, 1] - hps[:, -1, 1]
    h02 = hps[..., (7, 3), 1] - hps[..., (0, 4), 1]
    h13 = hps[..., (2, 6), 1] - hps[..., (1, 5), 1]
    cd = f * ph / (F.relu(center_h) * 4 + 1e-8)
    d02 = ((f * ph).unsqueeze(-1) / (F.relu(h02) * 4 + 1e-8)).mean(dim=1)
    d13 = ((f * ph).unsqueeze(-1) / (F.relu(h13) * 4 + 1e-8)).mean(dim=1)
    kpd = torch.clamp(torch.stack([cd, d02, d13], dim=-1), 0.1, 100)
    du = torch.clamp(gather("depth_uncertainty"), unc_range[0], unc_range[1])
    cu = torch.clamp(gather("corner_uncertainty"), unc_range[0], unc_range[1])
    unc = torch.cat((du, cu), dim=1).exp()
    depths = torch.cat((depth_decoded, kpd), dim=1)
    wts = 1 / unc
    wts = wts / wts.sum(dim=1, keepdim=True)
    merged = torch.sum(depths * wts, dim=1)
    mask = scores > score_thr
    rot = gather("rot")[mask]
    aidx = (rot[..., 1] > rot[..., 5]).float()
    alpha = (torch.atan(rot[..., 2] / rot[..., 3]) + (-0.5 * np.pi)) * aidx + (torch.atan(rot[..., 6] / rot[..., 7]) + (0.5 * np.pi)) * (1 - aidx)
    off = gather("reg")[mask]
    cx = (xs[mask] + off[..., 0]).unsqueeze(-1) * 4
    cy = (ys[mask] + off[..., 1]).unsqueeze(-1) * 4
    b2 = bbox2d[mask] * 4
    Hh, Ww = image_hw
    b2[:, 0] = torch.clamp(b2[:, 0], min=0); b2[:, 1] = torch.clamp(b2[:, 1], min=0)
    b2[:, 2] = torch.clamp(b2[:, 2], max=Ww); b2[:, 3] = torch.clamp(b2[:, 3], max=Hh)
    box = torch.cat([b2, cx, cy, merged[mask].unsqueeze(-1), dim[mask], alpha.unsqueeze(-1)], dim=1)
    sc = scores[mask]
    flat = (clses[mask].long() * height + ys[mask].long()) * width + xs[mask].long()
    keep = nms(box[:, :4], sc, nms_iou_thr)
    return sc[keep], box[keep], clses[mask].long()[keep], flat[keep]


def monoflex_forward(sd: SD, images, P2, cfg: dict, stages: dict | None = None):
    """MonoFlex.test_forward (R/detectors/KM3D.py:61-79), looped per image for the decode."""
    with torch.no_grad():
        ys = dla34(sd, "core.backbone", images)
        feat = dla_seg_upsample(sd, "core.deconv_layers", ys)
        names = list(cfg["head"]["layer_cfg"]["head_dict"].keys())
        outs = km3d_heads(sd, feat, names)
        if stages is not None:
            stages.update(features=feat, heads=outs, levels=ys)
        res = []
        for b in range(images.shape[0]):
            ob = {k: v[b:b + 1] for k, v in outs.items()}
            res.append(monoflex_get_bboxes(ob, P2[b:b + 1], images.shape[2:], cfg["head"]["test_cfg"].get("score_thr", 0.1),
                                           cfg["head"]["test_cfg"].get("nms_iou_thr", 0.5)))
        return res


def km3d_get_bboxes(output, P2, image_hw, score_thr=0.3, nms_iou_thr=0.5, K=100):
    """KM3DHead.get_bboxes + _decode (R/heads/km3d_head.py:155-314) + gen_position (R/utils/rtm3d_utils.py:314-455) for ONE
    image (maps [1, n, H, W]); the 1e-8 random jitter of A^T A (:447) is omitted."""
    from torchvision.ops import nms
    nmsf = lambda h: h * (F.max_pool2d(h, 3, stride=1, padding=1) == h).float()
    heat, hm_hp = nmsf(torch.sigmoid(output["hm"])), nmsf(torch.sigmoid(output["hm_hp"]))
    batch, cat, height, width = heat.shape
    HW = height * width
    ts, ti = torch.topk(heat.view(batch, cat, -1), K)
    ti = ti % HW
    scores, tind = torch.topk(ts.view(batch, -1), K)
    clses = (tind / K).int()
    inds = ti.view(batch, -1).gather(1, tind)
    ys, xs = (inds / width).int().float(), (inds % width).int().float()
    gat = lambda name, idx: output[name].permute(0, 2, 3, 1).reshape(batch, HW, -1).gather(
        1, idx.long().unsqueeze(2).expand(batch, idx.shape[1], output[name].shape[1]))
    J = 9
    kps = gat("hps", inds).view(batch, K, J * 2).clone()
    kps[..., ::2] += xs.view(batch, K, 1)
    kps[..., 1::2] += ys.view(batch, K, 1)
    reg = gat("reg", inds)
    xs2, ys2 = xs.view(batch, K, 1) + reg[:, :, 0:1], ys.view(batch, K, 1) + reg[:, :, 1:2]
    wh = gat("wh", inds)
    bboxes = torch.cat([xs2 - wh[..., 0:1] / 2, ys2 - wh[..., 1:2] / 2, xs2 + wh[..., 0:1] / 2, ys2 + wh[..., 1:2] / 2], dim=2)
    dim, rot = gat("dim", inds), gat("rot", inds)
    # keypoint refinement
    kpsj = kps.view(batch, K, J, 2).permute(0, 2, 1, 3).contiguous()
    reg_kps = kpsj.unsqueeze(3).expand(batch, J, K, K, 2)
    hs, hi = torch.topk(hm_hp.view(batch, J, -1), K)
    hi = hi % HW
    hys, hxs = (hi / width).int().float(), (hi % width).int().float()
    hpo = gat("hp_offset", hi.view(batch, -1)).view(batch, J, K, 2)
    hxs, hys = hxs + hpo[..., 0], hys + hpo[..., 1]
    m = (hs > 0.1).float()
    hs = (1 - m) * -1 + m * hs
    hys = (1 - m) * (-10000) + m * hys
    hxs = (1 - m) * (-10000) + m * hxs
    hm_kps = torch.stack([hxs, hys], dim=-1).unsqueeze(2).expand(batch, J, K, K, 2)
    dist = (((reg_kps - hm_kps) ** 2).sum(dim=4) ** 0.5)
    min_dist, min_ind = dist.min(dim=3)
    hsel = hs.gather(2, min_ind).unsqueeze(-1)
    min_dist = min_dist.unsqueeze(-1)
    hm_sel = hm_kps.gather(3, min_ind.view(batch, J, K, 1, 1).expand(batch, J, K, 1, 2)).view(batch, J, K, 2)
    ex = lambda t: t.view(batch, 1, K, 1).expand(batch, J, K, 1)
    l, t, r, b = ex(bboxes[:, :, 0]), ex(bboxes[:, :, 1]), ex(bboxes[:, :, 2]), ex(bboxes[:, :, 3])
    bad = (hm_sel[..., 0:1] < l) + (hm_sel[..., 0:1] > r) + (hm_sel[..., 1:2] < t) + (hm_sel[..., 1:2] > b) + \
          (hsel < 0.1) + (min_dist > (torch.max(b - t, r - l) * 0.3))
    bad = (bad > 0).float().expand(batch, J, K, 2)
    kps = ((1 - bad) * hm_sel + bad * kpsj).permute(0, 2, 1, 3).contiguous().view(batch, K, J * 2) * 4
    bboxes = bboxes * 4
    # gen_position
    calib = P2
    off_set = calib[:, 0, 3] / calib[:, 0, 0]
    si = torch.zeros_like(kps[:, :, 0:1]) + calib[:, 0:1, 0:1]
    aidx = (rot[:, :, 1] > rot[:, :, 5]).float()
    alpha = ((torch.atan(rot[:, :, 2] / rot[:, :, 3]) + (-0.5 * np.pi)) * aidx + (torch.atan(rot[:, :, 6] / rot[:, :, 7]) + (0.5 * np.pi)) * (1 - aidx)).unsqueeze(2)
    rot_y = alpha + torch.atan2(kps[:, :, 16:17] - calib[:, 0:1, 2:3], si)
    rot_y[rot_y > np.pi] -= 2 * np.pi
    rot_y[rot_y < -np.pi] += 2 * np.pi
    kpoint = kps[:, :, :16]
    f = calib[:, 0, 0].view(batch, 1, 1)
    cxy = torch.stack([calib[:, 0, 2], calib[:, 1, 2]], dim=-1).view(batch, 1, 2).repeat(1, 1, 8)
    kp_norm = (kpoint - cxy) / f
    ll, hh, ww = dim[:, :, 2:3], dim[:, :, 1:2], dim[:, :, 0:1]
    co, sn = torch.cos(rot_y), torch.sin(rot_y)
    lc, ls, wc, wsn, h2 = ll * 0.5 * co, ll * 0.5 * sn, ww * 0.5 * co, ww * 0.5 * sn, hh * 0.5
    Bx = [-lc - wsn, -lc + wsn, -lc + wsn, lc + wsn, lc + wsn, lc - wsn, lc - wsn, -lc - wsn]
    By = [-h2, -h2, h2, h2, -h2, -h2, h2, h2]
    Cc = [ls - wc, ls + wc, ls + wc, -ls + wc, -ls + wc, -ls - wc, -ls - wc, ls - wc]
    Bm = torch.cat([t for j in range(8) for t in (Bx[j], By[j])], dim=2)
    Cm = torch.cat([t for j in range(8) for t in (Cc[j], Cc[j])], dim=2)
    Bm = Bm - kp_norm * Cm
    const = torch.tensor([[-1.0, 0.0], [0.0, -1.0]] * 8).view(1, 1, 16, 2).expand(batch, K, -1, -1)
    A = torch.cat([const, kp_norm.unsqueeze(3)], dim=3).double().view(batch * K, 16, 3)
    AT = A.permute(0, 2, 1)
    pinv = torch.inverse(torch.bmm(AT, A))
    pos = torch.bmm(torch.bmm(pinv, AT).float(), Bm.view(batch * K, 16, 1).float()).view(batch, K, 3)
    pos[:, :, 0] -= off_set.unsqueeze(1)
    # get_bboxes
    sc, pos, bx, dm, al, cl = scores[0], pos[0], bboxes[0], dim[0], alpha[0], clses[0]
    mask = sc > score_thr
    p2 = P2[0]
    z3 = pos[mask][:, 2:3]
    cx3 = (pos[mask][:, 0:1] * p2[0, 0] + p2[0, 3] + p2[0, 2] * z3) / z3
    cy3 = (pos[mask][:, 1:2] * p2[1, 1] + p2[1, 3] + p2[1, 2] * z3) / z3
    b2 = bx[mask].clone()
    Hh, Ww = image_hw
    b2[:, 0] = torch.clamp(b2[:, 0], min=0); b2[:, 1] = torch.clamp(b2[:, 1], min=0)
    b2[:, 2] = torch.clamp(b2[:, 2], max=Ww); b2[:, 3] = torch.clamp(b2[:, 3], max=Hh)
    box = torch.cat([b2, cx3, cy3, z3, dm[mask], al[mask]], dim=1)
    flat = (cl[mask].long() * height + ys[0][mask].long()) * width + xs[0][mask].long()
    keep = nms(box[:, :4], sc[mask], nms_iou_thr)
    return sc[mask][keep], box[keep], cl[mask].long()[keep].unsqueeze(1), flat[keep]     # cls_indexes is [K, 1] in the reference (:276)


def km3d_forward(sd: SD, images, P2, cfg: dict, stages: dict | None = None):
    """KM3D.test_forward (R/detectors/KM3D.py:61-79), decode looped per image."""
    with torch.no_grad():
        ys = dla34(sd, "core.backbone", images)
        feat = dla_seg_upsample(sd, "core.deconv_layers", ys)
        outs = km3d_heads(sd, feat, list(cfg["head"]["layer_cfg"]["head_dict"].keys()))
        if stages is not None:
            stages.update(features=feat, heads=outs)
        return [km3d_get_bboxes({k: v[b:b + 1] for k, v in outs.items()}, P2[b:b + 1], images.shape[2:],
                                cfg["head"]["test_cfg"].get("score_thr", 0.1), cfg["head"]["test_cfg"].get("nms_iou_thr", 0.5))
                for b in range(images.shape[0])]
