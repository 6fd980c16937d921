"""-m gpu (runs last): the SURVEY 8(f) "next" rows on the GPU — `test_cfg.post_optimization` on the detector path, the CUDA form of the
hill climbing and of the input pipeline against their host forms (which tests/test_postopt_cpu.py / test_preprocess_cpu.py pin to the reference)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_detector_flag_applies_host_post_optimisation():
    """Yolo3D with head.test_cfg.post_optimization=True == the same forward without it, followed by postopt.post_process per image."""
    from visualdet3d_b200 import synth, postopt
    from visualdet3d_b200.detectors import build_synthetic_mono3d
    det, sd, cfg, _ = build_synthetic_mono3d("Yolo3D", seed=0)
    det = det.cuda().eval()
    img, P2 = synth.synth_mono_inputs(2, 96, 320, seed=4)
    with torch.no_grad():
        plain = det.forward_batch(img.cuda(), P2.cuda())
        det.post_optimization = True
        try:
            refined = det.forward_batch(img.cuda(), P2.cuda())
        finally:
            det.post_optimization = False
    for b, ((s, bx, c), (rs, rb, rc)) in enumerate(zip(plain, refined)):
        assert torch.equal(s, rs) and torch.equal(c, rc) and rb.device == bx.device
        want = postopt.post_process(bx, c, P2[b].numpy()) if len(s) else bx.cpu()
        assert torch.equal(rb.cpu(), want)


def test_device_hill_climbing_matches_host():
    """vd3d_post_opt (one thread per detection, in place on the fixed-capacity NMS layout) vs vd3d_post_opt_host on the same rows:
    same search, float64; the float32 alpha <-> yaw conversions use device atan2f instead of numpy's, so equality is to 1e-5 rad."""
    import os
    from conftest import GOLDEN
    from visualdet3d_b200 import postopt, _lib
    fx = np.load(os.path.join(GOLDEN, "postopt.npz"))
    P2 = fx["P2"]
    b = torch.from_numpy(fx["c2_in"])
    labels = torch.from_numpy(fx["c2_labels"]).long()
    want = postopt.post_process(b, labels, P2)
    K = b.shape[0]
    cap = 256
    boxes = torch.zeros(1, cap, 11)
    boxes[0, :K] = b
    cls = torch.zeros(1, cap, dtype=torch.int64)
    cls[0, :K] = labels
    boxes, cls = boxes.cuda(), cls.cuda()
    count = torch.tensor([K], dtype=torch.int32, device="cuda")
    P2d = torch.from_numpy(P2).view(1, 3, 4).cuda().contiguous()
    _lib.call("vd3d_post_opt", boxes.data_ptr(), cls.data_ptr(), count.data_ptr(), P2d.data_ptr(), 1, cap, 1280.0, 288.0, 0.4, 0.01, 3.0, 0,
              torch.cuda.current_stream().cuda_stream)
    got = boxes[0, :K].cpu()
    assert torch.equal(got[:, :10], want[:, :10])
    assert float(boxes[0, K:].abs().max()) == 0.0                       # rows beyond the count untouched
    d = (got[:, 10] - want[:, 10]).abs()
    d = torch.minimum(d, (d - 2 * np.pi).abs())
    print("device vs host hill climbing: max |d alpha|", float(d.max()), " rows within 1e-5:", int((d < 1e-5).sum()), "/", K)
    assert int((d < 1e-5).sum()) >= int(0.97 * K)


def test_device_input_pipeline_matches_host():
    """vd3d_preprocess (batched CUDA form, frames of two different sizes in one batch) vs vd3d_preprocess_host on the same frames."""
    import os
    import sys
    from conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    from make_golden_preprocess import frame
    from visualdet3d_b200 import preprocess as pp
    frames = [frame(0, 375, 1242), frame(1, 370, 1224)]
    want = np.stack([pp.preprocess_host(f, 100, (288, 1280)) for f in frames])
    got = pp.preprocess_batch(frames, 100, (288, 1280)).cpu().numpy()
    d = float(np.abs(got - want).max())
    print("device vs host input pipeline: max |diff|", d)
    assert got.shape == (2, 3, 288, 1280) and d < 1e-5
