"""-m gpu: MonoFlex (DLA-34 + DCNv2 up-sampling + CenterNet heads/decode) against the reference fixtures and the CPU oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_fixture, subsample_like
import torch_port as tp

pytestmark = pytest.mark.gpu


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def test_maxpool2x2_and_dw_convtranspose():
    from visualdet3d_b200 import engine as E
    from visualdet3d_b200._lib import call
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 24, 10, 14, generator=g)
    xa = E.Act(nhwc(x).cuda())
    o = E.Act(torch.empty(2, 5, 7, 24, device="cuda"))
    call("vd3d_maxpool2x2s2_nhwc", xa.ptr, 2, 10, 14, 24, 24, 0, o.ptr, 24, 0, None)
    assert torch.equal(o.to_nchw().cpu(), F.max_pool2d(x, 2, 2))
    for f in (2, 4):
        w = torch.randn(24, 1, 2 * f, 2 * f, generator=g)
        add = torch.randn(2, 24, 10 * f, 14 * f, generator=g)
        ref = F.conv_transpose2d(x, w, None, stride=f, padding=f // 2, groups=24) + add
        wk = w.reshape(24, -1).t().contiguous().cuda()
        aa = E.Act(nhwc(add).cuda())
        out = E.Act(torch.empty(2, 10 * f, 14 * f, 24, device="cuda"))
        call("vd3d_dw_convtranspose_nhwc", xa.ptr, 2, 10, 14, 24, 24, 0, wk.data_ptr(), f, aa.ptr, 24, 0, out.ptr, 24, 0, None)
        np.testing.assert_allclose(out.to_nchw().cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-5)


@pytest.fixture(scope="module")
def mf():
    from visualdet3d_b200.detectors import build_synthetic_monoflex
    det, sd, cfg = build_synthetic_monoflex(seed=0)
    return det.cuda().eval(), sd, cfg


def run_with_stages(det, img, P2):
    from visualdet3d_b200.engine import Act
    st = {}
    det.stage_hook = lambda name, v: st.__setitem__(name, v.to_nchw().cpu() if isinstance(v, Act) else v.detach().cpu().clone())
    try:
        with torch.no_grad():
            res = det.forward_batch(img.cuda(), P2.cuda())
    finally:
        det.stage_hook = None
    return res, st


def match_dets(got, ref, got_index, atol=1e-3):
    """same peak set; same order except between score-tied rows; values within atol (boxes rtol 1e-5 on top)."""
    s, bx, ci = [t.cpu() for t in got]
    rs, rb, rc, rflat = ref
    assert len(s) == len(rs), (len(s), len(rs))
    if len(s) == 0:
        return
    gi = got_index.cpu().long()
    assert torch.equal(torch.sort(gi)[0], torch.sort(rflat)[0]), "kept peak sets differ"
    if not torch.equal(gi, rflat):
        pos = {int(a): i for i, a in enumerate(rflat.tolist())}
        perm = torch.tensor([pos[int(a)] for a in gi.tolist()])
        for i in (perm != torch.arange(len(perm))).nonzero()[:, 0].tolist():
            assert abs(float(rs[perm[i]]) - float(rs[i])) < 1e-5
        rs, rb, rc = rs[perm], rb[perm], rc[perm]
    assert ci.shape == rc.shape and torch.equal(ci, rc)
    assert float((s - rs).abs().max()) < atol
    np.testing.assert_allclose(bx.numpy(), rb.numpy(), atol=atol, rtol=1e-5)


@pytest.mark.parametrize("tag", ["monoflex_96x320", "monoflex_192x640"])
def test_against_reference_fixture(mf, tag):
    from visualdet3d_b200 import synth
    det, sd, cfg = mf
    fx = load_fixture(tag)
    H, W, B, seed = [int(v) for v in fx["meta"]]
    img, P2 = synth.synth_mono_inputs(B, H, W, seed=1)
    res, st = run_with_stages(det, img, P2)
    rep = {"features": float(np.abs(subsample_like(st["features"], fx["features"]) - fx["features"]["samples"]).max())}
    off = det._plan["offsets"]
    for n, k in cfg["head"]["layer_cfg"]["head_dict"].items():
        got = st["heads"][:, off[n]:off[n] + k]
        rep[n] = float(np.abs(subsample_like(got, fx["head_" + n]) - fx["head_" + n]["samples"]).max())
    print(tag, "stage max|diff| vs reference:", rep)
    assert all(v < 1e-3 for v in rep.values()), rep
    ref = tp.monoflex_forward(sd, img, P2, cfg)
    for b in range(B):
        k = len(res[b][0])
        assert k == len(fx[f"scores_{b}"])
        match_dets(res[b], ref[b], det._last_decoder.anchor[b, :k])


def test_batch8_384x1280_properties(mf):
    """BASELINE configs[3] shape (DLA-34 + DCNv2, batch 8, 384x1280): batch invariance, determinism, one image vs the oracle."""
    from visualdet3d_b200 import synth
    det, sd, cfg = mf
    img, P2 = synth.synth_mono_inputs(8, 384, 1280, seed=9)
    ic, pc = img.cuda(), P2.cuda()
    with torch.no_grad():
        r1 = det.forward_batch(ic, pc)
        r2 = det.forward_batch(ic, pc)
        single = det([ic[2:3], pc[2:3]])
    assert all(torch.equal(x, y) for a, b in zip(r1, r2) for x, y in zip(a, b))
    assert all(torch.equal(x, y) for x, y in zip(r1[2], single))
    ref = tp.monoflex_forward(sd, img[2:3], P2[2:3], cfg)[0]
    k = len(single[0])
    match_dets(single, ref, det._last_decoder.anchor[0, :k])
    print("MonoFlex 384x1280 image: detections", k)


@pytest.mark.parametrize("tag", ["km3d_96x320", "km3d_192x640"])
def test_km3d_against_reference_fixture_and_oracle(tag):
    """KM3D: same network family, keypoint-refined least-squares decode (km3d_head.py:155-314, rtm3d_utils.py:314-455)."""
    from visualdet3d_b200 import synth
    from visualdet3d_b200.detectors import build_synthetic_monoflex
    det, sd, cfg = build_synthetic_monoflex(seed=0, name="KM3D")
    det = det.cuda().eval()
    fx = load_fixture(tag)
    H, W, B, seed = [int(v) for v in fx["meta"]]
    img, P2 = synth.synth_mono_inputs(B, H, W, seed=1)
    res, st = run_with_stages(det, img, P2)
    off = det._plan["offsets"]
    rep = {}
    for n, k in cfg["head"]["layer_cfg"]["head_dict"].items():
        got = st["heads"][:, off[n]:off[n] + k]
        rep[n] = float(np.abs(subsample_like(got, fx["head_" + n]) - fx["head_" + n]["samples"]).max())
    print(tag, "head max|diff| vs reference:", rep)
    assert all(v < 1e-3 for v in rep.values()), rep
    ref = tp.km3d_forward(sd, img, P2, cfg)
    for b in range(B):
        k = len(res[b][0])
        assert k == len(fx[f"scores_{b}"])
        match_dets(res[b], ref[b], det._last_decoder.anchor[b, :k], atol=5e-3)     # position = float64 3x3 solve of float32 keypoints
