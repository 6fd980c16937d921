"""Pins oracle/torch_port.py to the reference: the fixtures under tests/golden were produced by running the
UNMODIFIED reference (tests/golden/make_golden.py); here the port must reproduce them on CPU."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_fixture, subsample_like
from visualdet3d_b200 import synth
import torch_port as tp


def _stereo_setup(H, W, B, seed):
    shapes = json.load(open(os.path.join(GOLDEN, "stereo3d_keys.json")))
    sd = synth.synth_state_dict(shapes, seed)
    pm, ps = synth.synth_priors(16, 3, ["Car", "Pedestrian"])
    cfg = synth.stereo3d_cfg("/nonexistent")
    left, right, P2, P3 = synth.synth_stereo_inputs(B, H, W, seed=1)
    return sd, pm, ps, cfg, left, right, P2


@pytest.mark.parametrize("tag", ["stereo3d_96x320", "stereo3d_192x640"])
def test_stereo3d_port_matches_reference(tag):
    fx = load_fixture(tag)
    H, W, B, seed = [int(v) for v in fx["meta"]]
    sd, pm, ps, cfg, left, right, P2 = _stereo_setup(H, W, B, seed)
    st = {}
    outs = tp.stereo3d_forward(sd, left, right, P2, cfg, pm, ps, st)
    # stage tensors: the reference ran batch 1 per image, the port runs batched -> oneDNN may pick other kernels: 2e-4
    for nm in ["vol4", "vol8", "vol16", "features", "cls_preds", "reg_preds"]:
        got = subsample_like(st[nm], fx[nm])
        np.testing.assert_allclose(got, fx[nm]["samples"], rtol=0, atol=2e-4, err_msg=nm)
    bb = torch.cat([st["feat4"][:B], st["feat4"][B:]], 0)
    np.testing.assert_allclose(subsample_like(bb, fx["feat4"]), fx["feat4"]["samples"], atol=2e-4)
    np.testing.assert_array_equal(subsample_like(st["anchors"], fx["anchors"]), fx["anchors"]["samples"])
    np.testing.assert_array_equal(subsample_like(st["mean_std"], fx["mean_std"]), fx["mean_std"]["samples"])
    for b in range(B):
        np.testing.assert_array_equal(np.packbits(st["mask"][b].numpy()), fx[f"mask_{b}"])   # bit-exact useful mask
        s, bx, ci, _ = outs[b]
        assert len(s) == len(fx[f"scores_{b}"])
        np.testing.assert_array_equal(ci.numpy(), fx[f"cls_{b}"])                              # bit-exact class / keep set
        np.testing.assert_allclose(s.numpy(), fx[f"scores_{b}"], atol=1e-4)
        np.testing.assert_allclose(bx.numpy(), fx[f"bboxes_{b}"], atol=1e-3)


def test_psm_cosine_port_edge_cases():
    # narrower than the disparity range, single channel, batch > 1
    g = torch.Generator().manual_seed(0)
    L, R = torch.randn(2, 4, 3, 10, generator=g), torch.randn(2, 4, 3, 10, generator=g)
    c = tp.psm_cosine(L, R, 96, 4)
    assert c.shape == (2, 24, 3, 10)
    assert float(c[:, 10:].abs().max()) == 0.0           # planes i >= W stay zero
    assert float(c[:, 5, :, :5].abs().max()) == 0.0       # columns w < i stay zero
    np.testing.assert_allclose(c[:, 3, :, 3:], (L[..., 3:] * R[..., :-3]).mean(1))


@pytest.mark.parametrize("kind,tag", [("Yolo3D", "yolo3d_96x320"), ("Yolo3D", "yolo3d_288x1280"),
                                      ("GroundAwareYolo3D", "groundawareyolo3d_96x320"), ("GroundAwareYolo3D", "groundawareyolo3d_288x640")])
def test_mono3d_port_matches_reference(kind, tag):
    """Yolo3D (ResNet-18 + DCNv2 head, BASELINE configs[0] shape 1x3x288x1280) and GroundAwareYolo3D (ResNet-101 + LookGround)."""
    fx = load_fixture(tag)
    H, W, B, seed = [int(v) for v in fx["meta"]]
    shapes = json.load(open(os.path.join(GOLDEN, f"{kind.lower()}_keys.json")))
    sd = synth.synth_state_dict(shapes, seed, cls_gain=synth.CLS_GAIN.get(kind, 1.6))
    pm, ps = synth.synth_priors(16, 2, ["Car"])
    cfg = synth.mono3d_cfg("/nonexistent", kind)
    img, P2 = synth.synth_mono_inputs(B, H, W, seed=1)
    st = {}
    outs = tp.mono3d_forward(sd, img, P2, cfg, pm, ps, st)
    for nm in ["features", "cls_preds", "reg_preds"] + (["gac"] if "gac" in fx else []):
        np.testing.assert_allclose(subsample_like(st[nm], fx[nm]), fx[nm]["samples"], rtol=0, atol=2e-4, err_msg=nm)
    for b in range(B):
        np.testing.assert_array_equal(np.packbits(st["mask"][b].numpy()), fx[f"mask_{b}"])
        s, bx, ci, _ = outs[b]
        assert len(s) == len(fx[f"scores_{b}"])
        np.testing.assert_array_equal(ci.numpy(), fx[f"cls_{b}"])
        np.testing.assert_allclose(s.numpy(), fx[f"scores_{b}"], atol=1e-4)
        np.testing.assert_allclose(bx.numpy(), fx[f"bboxes_{b}"], atol=1e-3)


@pytest.mark.parametrize("tag", ["monoflex_96x320", "monoflex_192x640"])
def test_monoflex_port_matches_reference(tag):
    """MonoFlex (DLA-34 + 16 DCNv2 layers + 9 heads + CenterNet decode) vs the reference fixtures."""
    from visualdet3d_b200.detectors.centernet import monoflex_cfg
    fx = load_fixture(tag)
    H, W, B, seed = [int(v) for v in fx["meta"]]
    shapes = json.load(open(os.path.join(GOLDEN, "monoflex_keys.json")))
    sd = synth.synth_state_dict(shapes, seed)
    cfg = monoflex_cfg()
    img, P2 = synth.synth_mono_inputs(B, H, W, seed=1)
    st = {}
    outs = tp.monoflex_forward(sd, img, P2, cfg, st)
    np.testing.assert_allclose(subsample_like(st["features"], fx["features"]), fx["features"]["samples"], atol=2e-4)
    for n in cfg["head"]["layer_cfg"]["head_dict"]:
        np.testing.assert_allclose(subsample_like(st["heads"][n], fx["head_" + n]), fx["head_" + n]["samples"], atol=5e-4, err_msg=n)
    for b in range(B):
        s, bx, ci, _ = outs[b]
        assert len(s) == len(fx[f"scores_{b}"])
        np.testing.assert_array_equal(ci.numpy(), fx[f"cls_{b}"])
        np.testing.assert_allclose(s.numpy(), fx[f"scores_{b}"], atol=1e-4)
        np.testing.assert_allclose(bx.numpy(), fx[f"bboxes_{b}"], atol=1e-3, rtol=1e-5)


@pytest.mark.parametrize("tag", ["km3d_96x320", "km3d_192x640"])
def test_km3d_port_matches_reference(tag):
    """KM3D (DLA-34 + DCNv2 up-sampling + keypoint heads + least-squares position decode) vs the reference fixtures."""
    from visualdet3d_b200.detectors.centernet import km3d_cfg
    fx = load_fixture(tag)
    H, W, B, seed = [int(v) for v in fx["meta"]]
    shapes = json.load(open(os.path.join(GOLDEN, "km3d_keys.json")))
    sd = synth.synth_state_dict(shapes, seed)
    cfg = km3d_cfg()
    img, P2 = synth.synth_mono_inputs(B, H, W, seed=1)
    st = {}
    outs = tp.km3d_forward(sd, img, P2, cfg, st)
    for n in cfg["head"]["layer_cfg"]["head_dict"]:
        np.testing.assert_allclose(subsample_like(st["heads"][n], fx["head_" + n]), fx["head_" + n]["samples"], atol=5e-4, err_msg=n)
    for b in range(B):
        s, bx, ci, _ = outs[b]
        assert len(s) == len(fx[f"scores_{b}"])
        np.testing.assert_array_equal(ci.numpy(), fx[f"cls_{b}"])
        np.testing.assert_allclose(s.numpy(), fx[f"scores_{b}"], atol=1e-4)
        np.testing.assert_allclose(bx.numpy(), fx[f"bboxes_{b}"], atol=2e-3, rtol=1e-4)   # the reference jitters A^T A by 1e-8 randn
