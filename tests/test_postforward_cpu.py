"""SURVEY.md 8(f) rank 1 — post-forward geometry and the KITTI result writer (visualdet3d_b200/postforward.py) against fixtures
generated from the unmodified reference (tests/golden/make_golden_postforward.py): BackProjection, BBox3dProjector, the 2-D
rescale of `test_one` and write_result_to_file.  Float tensors bit-exact (same float32 ops in the same order), result files
character-identical."""
import os
import time

import numpy as np
import torch

from conftest import GOLDEN
from visualdet3d_b200 import postforward as pf

NAMES = ["Car", "Pedestrian", "Cyclist"]


def cases():
    fx = np.load(os.path.join(GOLDEN, "postforward.npz"))
    n = len({k.split("_")[0] for k in fx.files})
    for ci in range(n):
        yield ci, {k[len(f"c{ci}_"):]: fx[k] for k in fx.files if k.startswith(f"c{ci}_")}


def test_geometry_is_bit_exact():
    for ci, c in cases():
        bbox = torch.from_numpy(c["bbox"])
        b3 = pf.back_projection(bbox[:, 4:], c["P2"])
        assert np.array_equal(b3.numpy(), c["box3d"]), ci
        corners, homo, theta = pf.project_boxes(b3, c["P2"])
        assert np.array_equal(theta.numpy(), c["thetas"]), ci
        assert np.array_equal(corners.numpy(), c["corners"]) and np.array_equal(homo.numpy(), c["homo"]), ci
        assert np.array_equal(pf.alpha_to_theta(b3[:, 6], b3[:, 0], b3[:, 2], c["P2"]).numpy(), c["thetas"])
        assert np.array_equal(pf.rescale_boxes_2d(bbox[:, 0:4], c["P2"], c["oP"]).numpy(), c["box2d"]), ci
        assert np.array_equal(pf.rescale_boxes_2d_only(bbox[:, 0:4], 288, 375, 100).numpy(), c["box2d_only"]), ci
        assert np.array_equal(bbox.numpy(), c["bbox"])                    # inputs untouched


def test_kitti_result_text_is_identical(tmp_path):
    total, t0 = 0, time.perf_counter()
    for ci, c in cases():
        scores, bbox, cls = torch.from_numpy(c["scores"]), torch.from_numpy(c["bbox"]), torch.from_numpy(c["cls"])
        text = pf.detections_to_kitti(scores, bbox, cls, c["P2"], c["oP"], NAMES)
        assert text == bytes(c["text3d"]).decode(), ci
        total += len(scores)
        # 2-D-only branch: placeholders for the 3-D fields
        names = [NAMES[int(i)] for i in cls]
        t2 = pf.kitti_lines(scores.numpy(), pf.rescale_boxes_2d_only(bbox[:, 0:4], 288, 375, 100), obj_types=names)
        assert t2 == bytes(c["text2d"]).decode(), ci
        p = pf.write_result_file(str(tmp_path), 7 + ci, text)
        assert os.path.basename(p) == "%06d.txt" % (7 + ci) and open(p).read() == text
    dt = time.perf_counter() - t0
    print(f"post-forward + KITTI text for {total} detections (4 frames, both branches): {dt * 1e3:.1f} ms on the host")


def test_empty_and_threshold_behaviour():
    P2 = np.eye(3, 4, dtype=np.float32) * 700
    P2[2, 2] = 1
    e = pf.detections_to_kitti(torch.zeros(0), torch.zeros(0, 11), torch.zeros(0, dtype=torch.int64), P2, P2, NAMES)
    assert e == ""
    bbox = torch.tensor([[10., 20, 30, 40, 320, 200, 10, 1.6, 1.5, 3.9, 0.1]])
    assert pf.detections_to_kitti(torch.tensor([0.39]), bbox, torch.tensor([0]), P2, P2, NAMES) == ""          # below the 0.4 threshold
    line = pf.detections_to_kitti(torch.tensor([0.41]), bbox, torch.tensor([2]), P2, P2, NAMES)
    f = line.split()
    assert f[0] == "Cyclist" and len(f) == 16 and f[1] == "-1" and abs(float(f[-1]) - 0.41) < 1e-6
    assert abs(float(f[12]) - (200 * 10 / 700 + 0.5 * 1.5)) < 1e-5       # y = v z / fy (cy = ty = 0), moved to the box bottom: + h / 2
