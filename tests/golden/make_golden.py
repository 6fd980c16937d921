"""Golden-vector generator: runs the UNMODIFIED reference (imported from /root/reference, CPU, fp32) on the seeded
synthetic weights/inputs of visualdet3d_b200.synth and writes small fixtures next to this file.

    python tests/golden/make_golden.py [stereo3d] [yolo3d] [gac] [monoflex] [km3d]

Run in the build container only (the reference mount does not exist on the GPU box).  The fixtures pin
oracle/torch_port.py (tests/test_oracle_golden.py) and are compared against the CUDA path directly
(tests/test_*_gpu.py).  Large stage tensors are stored as a strided subsample plus two checksums.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import refload  # noqa: E402
from visualdet3d_b200 import synth  # noqa: E402

MAX_SAMPLES = 2048


def subsample(t: torch.Tensor) -> dict:
    """Fixture form of a stage tensor: strided samples + sum + abs-sum (float64)."""
    f = t.detach().reshape(-1).to(torch.float64)
    stride = max(1, f.numel() // MAX_SAMPLES)
    return dict(shape=np.array(t.shape, dtype=np.int64), stride=np.int64(stride),
                samples=f[::stride].to(torch.float32).numpy(), sum=np.float64(f.sum().item()),
                abssum=np.float64(f.abs().sum().item()))


def flatten_fixture(d: dict) -> dict:
    out = {}
    for k, v in d.items():
        if isinstance(v, dict):
            for kk, vv in v.items():
                out[f"{k}/{kk}"] = vv
        else:
            out[k] = v
    return out


def to_edict(d):
    from easydict import EasyDict
    if isinstance(d, dict):
        return EasyDict({k: to_edict(v) for k, v in d.items()})
    return d


def gen_stereo3d(H=96, W=320, B=2, seed=0, tag="stereo3d_96x320"):
    refload.load_reference()
    from visualDet3D.networks.utils.registry import DETECTOR_DICT
    obj_types = ["Car", "Pedestrian"]
    pm, ps = synth.synth_priors(16, 3, obj_types)
    tmp = tempfile.mkdtemp()
    synth.write_priors(tmp, pm, ps, obj_types)
    cfg = synth.stereo3d_cfg(tmp, obj_types)
    model = DETECTOR_DICT["Stereo3D"](to_edict(cfg))
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    with open(os.path.join(HERE, "stereo3d_keys.json"), "w") as f:
        json.dump({k: list(v) for k, v in shapes.items()}, f, indent=0)
    sd = synth.synth_state_dict(shapes, seed)
    missing = model.load_state_dict(sd, strict=False)
    print("missing (training-only buffers expected):", missing.missing_keys, "unexpected:", missing.unexpected_keys)
    model.eval()
    left, right, P2, P3 = synth.synth_stereo_inputs(B, H, W, seed=1)

    stages = {}
    hooks = []

    def cap(name):
        def fn(mod, inp, out):
            stages.setdefault(name, []).append(out.detach().clone() if torch.is_tensor(out) else out)
        return fn

    neck = model.core.neck
    hooks.append(neck.cost_volume_0.register_forward_hook(cap("vol4")))
    hooks.append(neck.cost_volume_1.register_forward_hook(cap("vol8")))
    hooks.append(neck.cost_volume_2.register_forward_hook(cap("vol16")))
    hooks.append(model.core.backbone.register_forward_hook(cap("backbone")))
    hooks.append(model.core.register_forward_hook(cap("core")))
    hooks.append(model.bbox_head.register_forward_hook(cap("head")))

    fix = {}
    outs = []
    with torch.no_grad():
        for b in range(B):  # reference asserts batch 1 (yolostereo3d_detector.py:78)
            s, bb, ci = model([left[b:b + 1], right[b:b + 1], P2[b:b + 1], P3[b:b + 1]])
            outs.append((s, bb, ci))
            # the anchors/mask the reference used for this image
            fix[f"mask_{b}"] = np.packbits(model.bbox_head.anchors.useful_mask[0].numpy())
    for h in hooks:
        h.remove()
    for b in range(B):
        s, bb, ci = outs[b]
        fix[f"scores_{b}"] = s.numpy()
        fix[f"bboxes_{b}"] = bb.numpy()
        fix[f"cls_{b}"] = ci.numpy()
        print(f"image {b}: {len(s)} detections; mask true = {int(model.bbox_head.anchors.useful_mask.sum())}")
    fix["anchors"] = subsample(model.bbox_head.anchors.anchors[0])
    fix["mean_std"] = subsample(model.bbox_head.anchors.anchor_mean_std)
    cat = lambda name, i=None: torch.cat([(x if i is None else x[i]) for x in stages[name]], dim=0)
    fix["vol4"] = subsample(cat("vol4"))
    fix["vol8"] = subsample(cat("vol8"))
    fix["vol16"] = subsample(cat("vol16"))
    # backbone sees [left; right] per image: regroup to the batched order [L0, L1, R0, R1]
    for j, nm in enumerate(["feat4", "feat8", "feat16"]):
        per = [x[j] for x in stages["backbone"]]
        fix[nm] = subsample(torch.cat([p[:1] for p in per] + [p[1:] for p in per], dim=0))
    fix["features"] = subsample(torch.cat([x["features"] for x in stages["core"]], dim=0))
    fix["cls_preds"] = subsample(cat("head", 0))
    fix["reg_preds"] = subsample(cat("head", 1))
    fix["meta"] = np.array([H, W, B, seed], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **flatten_fixture(fix))
    print("wrote", tag, {k: (v["shape"].tolist() if isinstance(v, dict) else np.asarray(v).shape) for k, v in fix.items()})

    # immediate pin of the oracle port (also asserted by tests/test_oracle_golden.py)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
    import torch_port as tp
    st = {}
    o = tp.stereo3d_forward(sd, left, right, P2, cfg, pm, ps, st)
    for b in range(B):
        print("oracle vs ref image", b, "n", len(o[b][0]), len(outs[b][0]),
              "max|dscore|", float((o[b][0] - outs[b][0]).abs().max()) if len(o[b][0]) == len(outs[b][0]) and len(o[b][0]) else None,
              "max|dbox|", float((o[b][1] - outs[b][1]).abs().max()) if len(o[b][0]) == len(outs[b][0]) and len(o[b][0]) else None)
    for nm in ["vol4", "vol8", "vol16", "features", "cls_preds", "reg_preds"]:
        ref = fix[nm]
        got = subsample(st[nm])
        print(nm, "max abs diff", float(np.abs(ref["samples"] - got["samples"]).max()), "ref absmean", ref["abssum"] / np.prod(ref["shape"]))


def gen_mono3d(kind, H=96, W=320, B=2, seed=0, depth=None):
    """Yolo3D (ResNet-18 + DCNv2 head; BASELINE configs[0]) / GroundAwareYolo3D (ResNet-101 + LookGround head; configs[2])."""
    refload.load_reference()
    from visualDet3D.networks.utils.registry import DETECTOR_DICT
    obj_types = ["Car"]
    pm, ps = synth.synth_priors(16, 2, obj_types)
    tmp = tempfile.mkdtemp()
    synth.write_priors(tmp, pm, ps, obj_types)
    cfg = synth.mono3d_cfg(tmp, kind, obj_types, depth)
    model = DETECTOR_DICT[kind](to_edict(cfg))
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    tag = f"{kind.lower()}_{H}x{W}"
    with open(os.path.join(HERE, f"{kind.lower()}_keys.json"), "w") as f:
        json.dump({k: list(v) for k, v in shapes.items()}, f, indent=0)
    sd = synth.synth_state_dict(shapes, seed, cls_gain=synth.CLS_GAIN.get(kind, 1.6))
    missing = model.load_state_dict(sd, strict=False)
    print("missing:", missing.missing_keys, "unexpected:", missing.unexpected_keys)
    model.eval()
    img, P2 = synth.synth_mono_inputs(B, H, W, seed=1)
    stages = {}
    hooks = [model.core.register_forward_hook(lambda m, i, o: stages.setdefault("features", []).append(o.detach().clone())),
             model.bbox_head.register_forward_hook(lambda m, i, o: stages.setdefault("head", []).append(o))]
    if kind == "GroundAwareYolo3D":
        hooks.append(model.bbox_head.reg_feature_extraction[0].register_forward_hook(
            lambda m, i, o: stages.setdefault("gac", []).append(o.detach().clone())))
    fix, outs = {}, []
    with torch.no_grad():
        for b in range(B):
            s, bb, ci = model([img[b:b + 1], P2[b:b + 1]])
            outs.append((s, bb, ci))
            fix[f"mask_{b}"] = np.packbits(model.bbox_head.anchors.useful_mask[0].numpy())
    for h in hooks:
        h.remove()
    for b in range(B):
        s, bb, ci = outs[b]
        fix[f"scores_{b}"], fix[f"bboxes_{b}"], fix[f"cls_{b}"] = s.numpy(), bb.numpy(), ci.numpy()
        print(f"{kind} image {b}: {len(s)} detections")
    fix["features"] = subsample(torch.cat(stages["features"], 0))
    fix["cls_preds"] = subsample(torch.cat([x[0] for x in stages["head"]], 0))
    fix["reg_preds"] = subsample(torch.cat([x[1] for x in stages["head"]], 0))
    if "gac" in stages:
        fix["gac"] = subsample(torch.cat(stages["gac"], 0))
    fix["meta"] = np.array([H, W, B, seed], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **flatten_fixture(fix))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
    import torch_port as tp
    st = {}
    o = tp.mono3d_forward(sd, img, P2, cfg, pm, ps, st)
    for b in range(B):
        same = len(o[b][0]) == len(outs[b][0])
        print("oracle vs ref image", b, "n", len(o[b][0]), len(outs[b][0]),
              "max|dbox|", float((o[b][1] - outs[b][1]).abs().max()) if same and len(o[b][0]) else None)
    for nm in ["features", "cls_preds", "reg_preds"] + (["gac"] if "gac" in fix else []):
        print(nm, "max abs diff", float(np.abs(fix[nm]["samples"] - subsample(st[nm])["samples"]).max()),
              "ref absmean", fix[nm]["abssum"] / np.prod(fix[nm]["shape"]))


def gen_monoflex(H=96, W=320, B=2, seed=0, kind="MonoFlex"):
    """MonoFlex / KM3D: DLA-34 + DCNv2 up-sampling + 9 heads + CenterNet-style decode (BASELINE configs[3] family)."""
    refload.load_reference()
    from visualDet3D.networks.utils.registry import DETECTOR_DICT
    from visualdet3d_b200.detectors.centernet import monoflex_cfg, km3d_cfg
    cfg = monoflex_cfg() if kind == "MonoFlex" else km3d_cfg()
    torch.manual_seed(0)
    model = DETECTOR_DICT[kind](to_edict(cfg))
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    with open(os.path.join(HERE, f"{kind.lower()}_keys.json"), "w") as f:
        json.dump({k: list(v) for k, v in shapes.items()}, f, indent=0)
    sd = synth.synth_state_dict(shapes, seed)
    missing = model.load_state_dict(sd, strict=False)
    print("missing:", missing.missing_keys, "unexpected:", missing.unexpected_keys)
    model.eval()
    img, P2 = synth.synth_mono_inputs(B, H, W, seed=1)
    stages = {}
    hooks = [model.core.register_forward_hook(lambda m, i, o: stages.setdefault("features", []).append(o.detach().clone())),
             model.bbox_head.register_forward_hook(lambda m, i, o: stages.setdefault("heads", []).append({k: v.detach().clone() for k, v in o.items()}))]
    fix, outs = {}, []
    with torch.no_grad():
        for b in range(B):
            outs.append(model([img[b:b + 1], P2[b:b + 1]]))
    for h in hooks:
        h.remove()
    for b in range(B):
        s, bb, ci = outs[b]
        fix[f"scores_{b}"], fix[f"bboxes_{b}"], fix[f"cls_{b}"] = s.numpy(), bb.numpy(), ci.numpy()
        print(f"{kind} image {b}: {len(s)} detections")
    fix["features"] = subsample(torch.cat(stages["features"], 0))
    for n in cfg["head"]["layer_cfg"]["head_dict"]:
        fix["head_" + n] = subsample(torch.cat([x[n] for x in stages["heads"]], 0))
    fix["meta"] = np.array([H, W, B, seed], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, f"{kind.lower()}_{H}x{W}.npz"), **flatten_fixture(fix))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
    import torch_port as tp
    st = {}
    o = (tp.monoflex_forward if kind == "MonoFlex" else tp.km3d_forward)(sd, img, P2, cfg, st)
    for b in range(B):
        same = len(o[b][0]) == len(outs[b][0])
        print("oracle vs ref image", b, "n", len(o[b][0]), len(outs[b][0]),
              "max|dscore|", float((o[b][0] - outs[b][0]).abs().max()) if same and len(o[b][0]) else None,
              "max|dbox|", float((o[b][1] - outs[b][1]).abs().max()) if same and len(o[b][0]) else None)
    print("features max abs diff", float(np.abs(fix["features"]["samples"] - subsample(st["features"])["samples"]).max()),
          "absmean", fix["features"]["abssum"] / np.prod(fix["features"]["shape"]))
    for n in cfg["head"]["layer_cfg"]["head_dict"]:
        print(n, "max abs diff", float(np.abs(fix["head_" + n]["samples"] - subsample(st["heads"][n])["samples"]).max()),
              "absmean", fix["head_" + n]["abssum"] / np.prod(fix["head_" + n]["shape"]))


if __name__ == "__main__":
    which = sys.argv[1:] or ["stereo3d"]
    torch.set_num_threads(os.cpu_count())
    if "stereo3d" in which:
        gen_stereo3d()
        gen_stereo3d(H=192, W=640, B=1, tag="stereo3d_192x640")
        gen_stereo3d(H=384, W=1280, B=1, tag="stereo3d_384x1280")      # BASELINE.json configs[1] shape
    if "yolo3d" in which:
        gen_mono3d("Yolo3D", 96, 320, 2)
        gen_mono3d("Yolo3D", 288, 1280, 1)       # BASELINE.json configs[0]
    if "km3d" in which:
        gen_monoflex(96, 320, 2, kind="KM3D")
        gen_monoflex(192, 640, 1, kind="KM3D")
        gen_monoflex(384, 1280, 1, kind="KM3D")     # BASELINE.json configs[3] shape
    if "monoflex" in which:
        gen_monoflex(96, 320, 2)
        gen_monoflex(192, 640, 1)
        gen_monoflex(384, 1280, 1)                  # BASELINE.json configs[3] shape
    if "gac" in which:
        gen_mono3d("GroundAwareYolo3D", 96, 320, 2)
        gen_mono3d("GroundAwareYolo3D", 288, 640, 1)
        gen_mono3d("GroundAwareYolo3D", 288, 1280, 1)   # BASELINE.json configs[2] shape
