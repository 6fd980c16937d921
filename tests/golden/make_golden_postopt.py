"""Golden vectors for visualdet3d_b200/postopt.py from the UNMODIFIED reference's numba hill climbing (build container only):
`post_opt` (networks/lib/fast_utils/hill_climbing.py:7-22) applied like `_post_process` (heads/detection_3d_head.py:294-308) on seeded
synthetic detections whose 2-D boxes are the (perturbed) hulls of plausible 3-D boxes.   python tests/golden/make_golden_postopt.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import refload  # noqa: E402

P2 = np.array([[5.02790613e+02, 0.0, 4.29568996e+02, 3.25392427e+01], [0.0, 5.02790613e+02, 5.72491378e+01, -5.99834524e-01],
               [0.0, 0.0, 1.0, 4.98101600e-03]], dtype=np.float32)          # the calibration of the reference's embedded example (hill_climbing.py:128-131)


def synth(seed, K):
    rng = np.random.RandomState(seed)
    z = rng.uniform(1.5, 60, K).astype(np.float32)
    x3 = rng.uniform(-0.6, 0.6, K) * z
    y3 = rng.uniform(0.8, 2.0, K)
    whl = np.stack([rng.uniform(1.4, 2.0, K), rng.uniform(1.3, 1.9, K), rng.uniform(3.0, 5.0, K)], 1).astype(np.float32)
    cx = (P2[0, 0] * x3 + P2[0, 2] * z + P2[0, 3]) / z
    cy = (P2[1, 1] * y3 + P2[1, 2] * z + P2[1, 3]) / z
    alpha = rng.uniform(-3.1, 3.1, K).astype(np.float32)
    half_w = P2[0, 0] * (whl[:, 2] * 0.45 + rng.uniform(-0.3, 0.3, K)) / z
    half_h = P2[1, 1] * (whl[:, 1] * 0.5 + rng.uniform(-0.1, 0.1, K)) / z
    box = np.stack([cx - half_w, cy - half_h, cx + half_w, cy + half_h], 1) + rng.uniform(-2, 2, (K, 4))
    bbox = np.concatenate([box, cx[:, None], cy[:, None], z[:, None], whl, alpha[:, None]], 1).astype(np.float32)
    labels = rng.randint(0, 2, K)
    return torch.from_numpy(bbox), torch.from_numpy(labels)


def main():
    refload.load_reference()
    from visualDet3D.networks.lib.fast_utils.hill_climbing import post_opt
    from visualDet3D.networks.utils.utils import BackProjection
    out = {"P2": P2}
    for ci, (seed, K) in enumerate([(0, 96), (1, 5), (2, 160)]):
        bboxes, labels = synth(seed, K)
        bbox2d, bbox3d = bboxes[:, 0:4].clone(), bboxes[:, 4:].clone()
        state = BackProjection().forward(bbox3d, P2)
        n = 0
        for i in range(K):
            if state[i, 2] > 3 and labels[i] == 0:
                bbox3d[i] = post_opt(bbox2d[i], state[i], P2, bbox3d[i, 0].item(), bbox3d[i, 1].item())
                n += 1
        ref = torch.cat([bbox2d, bbox3d], dim=-1)
        out.update({f"c{ci}_in": bboxes.numpy(), f"c{ci}_labels": labels.numpy(), f"c{ci}_out": ref.numpy()})
        print(f"case {ci}: K={K}, refined {n}, changed alpha on {(ref[:, 10] != bboxes[:, 10]).sum().item()} rows")
    np.savez_compressed(os.path.join(HERE, "postopt.npz"), **out)


if __name__ == "__main__":
    main()
