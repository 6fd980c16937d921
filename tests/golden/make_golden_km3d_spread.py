"""Run-to-run spread of the UNMODIFIED reference KM3D decode.

`gen_position` (R/networks/utils/rtm3d_utils.py:439-449) adds `torch.randn_like(pinv) * 1e-8` to the float64 normal matrix
A^T A before inverting it, so the reference's own outputs depend on the state of torch's global RNG.  This script runs the
reference forward N_SEEDS times on the same seeded weights / inputs with a different `torch.manual_seed` in front of every
forward and stores, per detection row, the min and max of every output column -> tests/golden/km3d_spread.npz.

    python tests/golden/make_golden_km3d_spread.py

The GPU parity test (tests/test_monoflex_gpu.py) holds every column to 1e-3 EXCEPT where the reference disagrees with
itself: there the tolerance is 1e-3 + the reference's own spread of that entry.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import refload  # noqa: E402
from make_golden import to_edict  # noqa: E402
from visualdet3d_b200 import synth  # noqa: E402

N_SEEDS = 8
CASES = [(96, 320, 2), (192, 640, 1), (384, 1280, 1)]


def main():
    refload.load_reference()
    from visualDet3D.networks.utils.registry import DETECTOR_DICT
    from visualdet3d_b200.detectors.centernet import km3d_cfg
    cfg = km3d_cfg()
    torch.manual_seed(0)
    model = DETECTOR_DICT["KM3D"](to_edict(cfg))
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synth.synth_state_dict(shapes, 0), strict=False)
    model.eval()
    out = {"n_seeds": np.int64(N_SEEDS)}
    for (H, W, B) in CASES:
        img, P2 = synth.synth_mono_inputs(B, H, W, seed=1)
        for b in range(B):
            runs = []
            for s in range(N_SEEDS):
                torch.manual_seed(1000 + s)
                with torch.no_grad():
                    sc, bb, ci = model([img[b:b + 1], P2[b:b + 1]])
                runs.append((sc.numpy().copy(), bb.numpy().copy(), ci.numpy().copy()))
            assert all(np.array_equal(r[2], runs[0][2]) and np.array_equal(r[0], runs[0][0]) for r in runs), "scores / classes must not depend on the jitter"
            bbs = np.stack([r[1] for r in runs]).astype(np.float64)       # [S, K, 11]
            lo, hi = bbs.min(0), bbs.max(0)
            tag = f"{H}x{W}_{b}"
            out[f"{tag}/min"], out[f"{tag}/max"] = lo.astype(np.float32), hi.astype(np.float32)
            out[f"{tag}/scores"] = runs[0][0]
            spread = hi - lo
            print(f"KM3D {tag}: K={bbs.shape[1]}  max spread per column:", np.array2string(spread.max(0), precision=2),
                  " rows with spread > 1e-3:", int((spread.max(1) > 1e-3).sum()))
    np.savez_compressed(os.path.join(HERE, "km3d_spread.npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    main()
