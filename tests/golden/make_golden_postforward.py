"""Golden vectors for visualdet3d_b200/postforward.py from the UNMODIFIED reference (build container only):
BackProjection / BBox3dProjector (networks/utils/utils.py:198-278), the 2-D rescale of test_one (pipelines/evaluators.py:118-127)
and write_result_to_file (data/kitti/utils.py:162-201), on seeded synthetic detections.   python tests/golden/make_golden_postforward.py"""
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


def synth_case(seed: int, K: int):
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(K, generator=g) * 1280
    v = torch.rand(K, generator=g) * 288 + 40
    z = torch.rand(K, generator=g) * 60 + 2
    whl = torch.rand(K, 3, generator=g) * 2 + 0.5
    alpha = (torch.rand(K, generator=g) - 0.5) * 6.2
    x1 = u - torch.rand(K, generator=g) * 80; y1 = v - torch.rand(K, generator=g) * 40
    x2 = u + torch.rand(K, generator=g) * 80; y2 = v + torch.rand(K, generator=g) * 40
    bbox = torch.stack([x1, y1, x2, y2, u, v, z, whl[:, 0], whl[:, 1], whl[:, 2], alpha], dim=1)
    scores = torch.rand(K, generator=g) * 0.6 + 0.3                 # some fall below the 0.4 write threshold
    cls = torch.randint(0, 3, (K,), generator=g)
    P2 = np.array([[707.0493 * 0.9, 0.0, 604.0814 * 0.9, 45.75831 * 0.9], [0.0, 707.0493 * 0.9, 180.5066 * 0.9 - 90.0, -0.3454157 * 0.9], [0.0, 0.0, 1.0, 0.004981016]], dtype=np.float32)
    oP = np.array([[707.0493, 0.0, 604.0814, 45.75831], [0.0, 707.0493, 180.5066, -0.3454157], [0.0, 0.0, 1.0, 0.004981016]], dtype=np.float32)
    return scores, bbox, cls, P2, oP


def main():
    refload.load_reference()
    from visualDet3D.networks.utils.utils import BackProjection, BBox3dProjector
    from visualDet3D.data.kitti.utils import write_result_to_file
    names = ["Car", "Pedestrian", "Cyclist"]
    out = {}
    for ci, (seed, K) in enumerate([(0, 37), (1, 1), (2, 0), (3, 200)]):
        scores, bbox, cls, P2, oP = synth_case(seed, K)
        bbox_2d = bbox[:, 0:4].clone()
        state = bbox[:, 4:].clone()
        b3 = BackProjection()(state, P2)
        corners, homo, thetas = BBox3dProjector()(b3, b3.new(P2))
        sx, sy = oP[0, 0] / P2[0, 0], oP[1, 1] / P2[1, 1]
        bbox_2d[:, 0:4:2] += oP[0, 2] / sx - P2[0, 2]
        bbox_2d[:, 1:4:2] += oP[1, 2] / sy - P2[1, 2]
        bbox_2d[:, 0:4:2] *= sx
        bbox_2d[:, 1:4:2] *= sy
        d = tempfile.mkdtemp()
        obj = [names[int(i)] for i in cls]
        write_result_to_file(d, 7, scores, bbox_2d, b3.clone(), thetas, obj)
        text3d = open(os.path.join(d, "000007.txt")).read()
        b2o = bbox[:, 0:4].clone()
        b2o[:, 0:4] *= (375 - 100) / 288
        b2o[:, 1:4:2] += 100
        write_result_to_file(d, 8, scores.numpy(), b2o, obj_types=obj)
        text2d = open(os.path.join(d, "000008.txt")).read()
        p = f"c{ci}_"
        out.update({p + "scores": scores.numpy(), p + "bbox": bbox.numpy(), p + "cls": cls.numpy(), p + "P2": P2, p + "oP": oP,
                    p + "box3d": b3.numpy(), p + "corners": corners.numpy(), p + "homo": homo.numpy(), p + "thetas": thetas.numpy(),
                    p + "box2d": bbox_2d.numpy(), p + "box2d_only": b2o.numpy(),
                    p + "text3d": np.frombuffer(text3d.encode(), dtype=np.uint8), p + "text2d": np.frombuffer(text2d.encode(), dtype=np.uint8)})
        print(f"case {ci}: K={K}, {text3d.count(chr(10))} lines written")
    np.savez_compressed(os.path.join(HERE, "postforward.npz"), **out)


if __name__ == "__main__":
    main()
