"""Run single tensor-core convs of the two dominant shapes (for ncu captures / quick timing).
usage: python tools/prof_conv.py [reps]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visualdet3d_b200 import engine as E

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
SHAPES = [("head 1408->1408 @24x80 B8", 8, 24, 80, 1408, 1408), ("layer1 64->64 @96x320 B16", 16, 96, 320, 64, 64),
          ("layer2 128->128 @48x160 B16", 16, 48, 160, 128, 128), ("layer3 256->256 @24x80 B16", 16, 24, 80, 256, 256)]
g = torch.Generator().manual_seed(0)
for name, B, H, W, Cin, Cout in SHAPES:
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / np.sqrt(Cin * 9)
    layer = E.ConvLayer(w, None, None, pad=1, relu=True, device="cuda", engine="tc16")
    x = E.Act(torch.randn(B, H, W, Cin, generator=g).cuda(), 0, None, torch.zeros(2, B, H, W, Cin, device="cuda", dtype=torch.float16))
    E.split_lo(x)
    out = E.Act(torch.empty(B, H, W, Cout, device="cuda"), 0, None, torch.zeros(2, B, H, W, Cout, device="cuda", dtype=torch.float16))
    for mode in os.environ.get("PROF_MODES", "0,2,1").split(","):
        os.environ["VD3D_TC_HALO"] = mode
        layer(x, out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            layer(x, out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        fl = 2.0 * B * H * W * Cin * Cout * 9
        print(f"{name}  halo={mode}  {ms*1e3:8.1f} us   {fl/ms/1e9:7.1f} TFLOP/s (x3 passes = {3*fl/ms/1e9:7.1f})", flush=True)
