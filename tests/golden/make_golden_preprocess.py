"""Golden vectors for visualdet3d_b200/preprocess.py from the UNMODIFIED reference augmentation classes (build container only):
ConvertToFloat -> CropTop(100) -> Resize((288, 1280)) -> Normalize (config/Stereo3D_example:102-107) on seeded random uint8 frames of
KITTI sizes, with P2 / P3.   python tests/golden/make_golden_preprocess.py"""
import os
import sys
from copy import deepcopy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import refload  # noqa: E402


def frame(seed, H, W):
    rng = np.random.RandomState(seed)
    base = rng.randint(0, 256, (H // 8 + 2, W // 8 + 2, 3)).astype(np.float32)        # smooth-ish content plus noise (resize is tested on both)
    img = np.kron(base, np.ones((8, 8, 1), dtype=np.float32))[:H, :W] * 0.7 + rng.randint(0, 77, (H, W, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    refload.load_reference()
    from visualDet3D.data.pipeline.stereo_augmentator import ConvertToFloat, CropTop, Resize, Normalize
    mean, std = np.array([0.485, 0.456, 0.406]), np.array([0.229, 0.224, 0.225])
    P2 = np.array([[7.215377e+02, 0.0, 6.095593e+02, 4.485728e+01], [0.0, 7.215377e+02, 1.728540e+02, 2.163791e-01], [0.0, 0.0, 1.0, 2.745884e-03]])
    P3 = np.array([[7.215377e+02, 0.0, 6.095593e+02, -3.395242e+02], [0.0, 7.215377e+02, 1.728540e+02, 2.199936e+00], [0.0, 0.0, 1.0, 2.729905e-03]])
    out = {"P2": P2, "P3": P3}
    for ci, (seed, H, W, crop, size) in enumerate([(0, 375, 1242, 100, (288, 1280)), (1, 370, 1224, 100, (288, 1280)), (2, 376, 1241, 88, (288, 1280)),
                                                   (3, 200, 640, 40, (96, 320)), (4, 375, 1242, 100, (288, 1400))]):
        l, r = frame(seed, H, W), frame(seed + 100, H, W)
        p2, p3 = deepcopy(P2), deepcopy(P3)
        data = (l, r, p2, p3, None, None, None)
        for aug in (ConvertToFloat(), CropTop(crop_top_index=crop), Resize(size=size), Normalize(mean=mean, stds=std)):
            data = aug(*data)
        lo, ro, p2o, p3o = data[0], data[1], data[2], data[3]
        lo, ro = lo.transpose(2, 0, 1), ro.transpose(2, 0, 1)                 # collate_fn: [H, W, 3] -> [3, H, W]
        st = max(1, lo.size // 4096)
        out.update({f"c{ci}_meta": np.array([seed, H, W, crop, size[0], size[1]]), f"c{ci}_P2": p2o, f"c{ci}_P3": p3o,
                    f"c{ci}_l_samples": lo.reshape(-1)[::st].astype(np.float32), f"c{ci}_r_samples": ro.reshape(-1)[::st].astype(np.float32),
                    f"c{ci}_stride": np.int64(st), f"c{ci}_l_sum": np.float64(lo.astype(np.float64).sum()), f"c{ci}_l_abssum": np.float64(np.abs(lo.astype(np.float64)).sum()),
                    f"c{ci}_last_col": lo[:, :, -1].astype(np.float32), f"c{ci}_first_row": lo[:, 0, :].astype(np.float32)})
        print(f"case {ci}: {H}x{W} crop {crop} -> {lo.shape}")
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)


if __name__ == "__main__":
    main()
