"""Worker of tests/test_reference_seam.py::test_reference_modules_and_pipelines_over_b200_ops (own process, GPU box).

The integration seam, executed: the reference's UNMODIFIED Python (oracle/_ref/visualDet3D, a verbatim copy made by oracle/build_ref.py)
runs on the GPU with `visualdet3d_b200.ops.dcn` / `.ops.iou3d` standing in for its pybind modules `deform_conv_ext` / `iou3d_cuda`
(one `sys.modules` assignment each, oracle/refload.py::_install_ext), its unmodified test pipelines drive the B200 detector classes
installed by `plugin.install_into_reference()`, and its own unmodified Stereo3D detector is run on the same GPU as a second oracle.
Prints one JSON line."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import build_ref  # noqa: E402
import refload  # noqa: E402


def load_fixture(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return {k: z[k] for k in z.files}


def main():
    torch.backends.cudnn.allow_tf32 = False                  # the comparisons below are fp32-grade
    torch.backends.cuda.matmul.allow_tf32 = False
    ref_dcn, ref_iou = build_ref.load("ref_deform_conv_ext"), build_ref.load("ref_iou3d_cuda")
    from visualdet3d_b200.ops import dcn as our_dcn, iou3d as our_iou
    from visualdet3d_b200 import plugin, synth
    import visualdet3d_b200.detectors  # noqa: F401
    refload.load_reference(device="cuda", dcn_ext=our_dcn, iou3d_ext=our_iou)
    out = {}
    # ---- 1. the reference's unmodified deform_conv.py on our extension module -------------------------------------------------------
    from visualDet3D.networks.lib.ops import ModulatedDeformConvPack, DeformConvPack
    from visualDet3D.networks.lib.ops.dcn import deform_conv as dc
    assert dc.deform_conv_ext is our_dcn
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 64, 24, 40, generator=g).cuda()
    for key, cls in (("dcn_v2_rel_err", ModulatedDeformConvPack), ("dcn_v1_rel_err", DeformConvPack)):
        m = cls(64, 96, 3, stride=1, padding=1).cuda()
        m.weight.data = (torch.randn(m.weight.shape, generator=g) * 0.05).cuda()
        m.conv_offset.weight.data = (torch.randn(m.conv_offset.weight.shape, generator=g) * 0.02).cuda()
        m.conv_offset.bias.data = (torch.randn(m.conv_offset.bias.shape, generator=g) * 0.5).cuda()
        with torch.no_grad():
            dc.deform_conv_ext = our_dcn
            y_ours = m(x).clone()
            dc.deform_conv_ext = ref_dcn                      # the reference's own compiled extension under the same Python
            y_ref = m(x).clone()
            dc.deform_conv_ext = our_dcn
        out[key] = float((y_ours - y_ref).abs().max() / y_ref.abs().max())
        # ... and its unmodified autograd Function (deform_conv.py:55-230) training through the B200 backward entries
        grads = []
        for ext in (our_dcn, ref_dcn):
            dc.deform_conv_ext = ext
            xg = x.clone().requires_grad_(True)
            m.zero_grad()
            (m(xg) * torch.linspace(-1, 1, 96, device="cuda").view(1, -1, 1, 1)).sum().backward()
            grads.append([xg.grad.clone(), m.weight.grad.clone(), m.conv_offset.weight.grad.clone()])
        dc.deform_conv_ext = our_dcn
        out[key.replace("rel_err", "grad_rel_err")] = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(*grads))
    # ---- 2. the reference's unmodified iou3d.py on our extension module ---------------------------------------------------------------
    from visualDet3D.networks.lib.ops.iou3d import iou3d as ri
    a = torch.rand(40, 7, generator=g) * torch.tensor([20, 2, 40, 1, 1, 3, 6.28]) + torch.tensor([-10, 0, 2, 1.2, 1.4, 3.0, -3.14])
    b = a[torch.randperm(40, generator=g)] + torch.randn(40, 7, generator=g) * 0.3
    a, b = a.cuda(), b.cuda()
    assert ri.boxes_overlap_bev_gpu is our_iou.boxes_overlap_bev_gpu
    iou_ours = ri.boxes_iou3d_gpu(a, b).clone()
    bev = ri.boxes3d_to_bev_torch(a).contiguous()
    order = torch.argsort(torch.rand(40, generator=g).cuda(), descending=True)
    keep_o = torch.zeros(40, dtype=torch.int64)
    n_o = our_iou.nms_gpu(bev[order].contiguous(), keep_o, 0.1)
    for nm in ("boxes_iou_bev_gpu", "boxes_overlap_bev_gpu", "nms_normal_gpu", "nms_gpu"):
        setattr(ri, nm, getattr(ref_iou, nm))
    iou_ref = ri.boxes_iou3d_gpu(a, b).clone()
    keep_r = torch.zeros(40, dtype=torch.int64)
    n_r = ref_iou.nms_gpu(bev[order].contiguous(), keep_r, 0.1)
    out["iou3d_max_err"] = float((iou_ours - iou_ref).abs().max())
    out["iou3d_nms_equal"] = bool(n_o == n_r and torch.equal(keep_o[:n_o], keep_r[:n_r]))
    # ---- 3. the reference's unmodified test pipelines driving the B200 detectors ----------------------------------------------------------
    from visualDet3D.networks.utils import registry as ref_registry
    RefStereo3D = ref_registry.DETECTOR_DICT["Stereo3D"]
    ref = plugin.install_into_reference()
    tmp = tempfile.mkdtemp()
    obj = ["Car", "Pedestrian"]
    pm, ps = synth.synth_priors(16, 3, obj)
    synth.write_priors(tmp, pm, ps, obj)
    cfg = refload.to_edict(dict(obj_types=obj, detector=synth.stereo3d_cfg(tmp, obj)))
    det = ref.DETECTOR_DICT[cfg.detector.name](cfg.detector)                     # scripts/eval.py:37
    shapes = {k: tuple(v.shape) for k, v in det.state_dict().items()}
    sd = synth.synth_state_dict(shapes, 0)
    det.load_state_dict(sd, strict=False)                                        # scripts/eval.py:42
    det = det.cuda()
    det.eval()
    fx = load_fixture("stereo3d_96x320")
    left, right, P2, P3 = synth.synth_stereo_inputs(2, 96, 320, seed=1)
    data = [left[:1], right[:1], P2[:1].numpy(), P3[:1].numpy()]                 # what collate_fn hands to the test function
    test_fn = ref.PIPELINE_DICT["test_stereo_detection"]                         # evaluators.py looks it up by cfg.trainer.test_func
    scores, bbox, names = test_fn(data, det, None, cfg=cfg)
    out["stereo"] = dict(count=len(scores), fixture_count=len(fx["scores_0"]),
                         names_ok=names == [obj[int(i)] for i in fx["cls_0"]],
                         max_score_diff=float(np.abs(scores.cpu().numpy() - fx["scores_0"]).max()) if len(scores) == len(fx["scores_0"]) else 1e9,
                         max_box_diff=float(np.abs(bbox.cpu().numpy() - fx["bboxes_0"]).max()) if len(scores) == len(fx["scores_0"]) else 1e9)
    # mono: Yolo3D (DCNv2 head) through test_mono_detection
    pm1, ps1 = synth.synth_priors(16, 2, ["Car"])
    tmp1 = tempfile.mkdtemp()
    synth.write_priors(tmp1, pm1, ps1, ["Car"])
    mcfg = refload.to_edict(dict(obj_types=["Car"], detector=synth.mono3d_cfg(tmp1, "Yolo3D", ["Car"], None)))
    mdet = ref.DETECTOR_DICT[mcfg.detector.name](mcfg.detector)
    msd = synth.synth_state_dict({k: tuple(v.shape) for k, v in mdet.state_dict().items()}, 0, cls_gain=synth.CLS_GAIN.get("Yolo3D", 1.6))
    mdet.load_state_dict(msd, strict=False)
    mdet = mdet.cuda().eval()
    mfx = load_fixture("yolo3d_96x320")
    img, mP2 = synth.synth_mono_inputs(2, 96, 320, seed=1)
    s2, b2, n2 = ref.PIPELINE_DICT["test_mono_detection"]([img[:1], mP2[:1].numpy()], mdet, None, cfg=mcfg)
    same = len(s2) == len(mfx["scores_0"])
    # rows can swap between score-tied detections: align on the score order of the fixture through a stable sort of both
    out["mono"] = dict(count=len(s2), fixture_count=len(mfx["scores_0"]), names_ok=n2 == ["Car"] * len(s2),
                       max_score_diff=float(np.abs(s2.cpu().numpy() - mfx["scores_0"]).max()) if same else 1e9,
                       max_box_diff=float(np.abs(b2.cpu().numpy() - mfx["bboxes_0"]).max()) if same else 1e9)
    # ---- 4. the UNMODIFIED reference Stereo3D on this GPU (its own CUDA code path) vs the B200 class --------------------------------------
    rdet = RefStereo3D(cfg.detector)
    rdet.load_state_dict(sd, strict=False)
    rdet = rdet.cuda().eval()
    with torch.no_grad():
        rs, rb, rc = rdet([left[1:2].cuda(), right[1:2].cuda(), P2[1:2].cuda(), P3[1:2].cuda()])
        os_, ob, oc = det([left[1:2].cuda(), right[1:2].cuda(), P2[1:2].cuda(), P3[1:2].cuda()])
    same = len(rs) == len(os_)
    out["ref_gpu"] = dict(count=len(rs), b200_count=len(os_), cls_equal=bool(same and torch.equal(rc, oc)),
                          max_score_diff=float((rs - os_).abs().max()) if same and len(rs) else (0.0 if same else 1e9),
                          max_box_diff=float((rb - ob).abs().max()) if same and len(rs) else (0.0 if same else 1e9))
    print("SEAM_JSON " + json.dumps(out))


if __name__ == "__main__":
    main()
