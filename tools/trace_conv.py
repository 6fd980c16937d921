"""Pipeline timeline of the persistent tensor-core conv (CTA 0): per k-block clock64 stamps -> load latency, MMA-thread wait time, period.
usage: python tools/trace_conv.py [shape-index ...]   (shapes of tools/prof_conv.py)"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visualdet3d_b200 import engine as E, _lib

SHAPES = [("head 1408->1408 @24x80 B8", 8, 24, 80, 1408, 1408), ("layer1 64->64 @96x320 B16", 16, 96, 320, 64, 64),
          ("layer2 128->128 @48x160 B16", 16, 48, 160, 128, 128), ("layer3 256->256 @24x80 B16", 16, 24, 80, 256, 256)]
sel = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3]
N = 600
g = torch.Generator().manual_seed(0)
lib = _lib.load()
for si in sel:
    name, B, H, W, Cin, Cout = SHAPES[si]
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / np.sqrt(Cin * 9)
    layer = E.ConvLayer(w, None, None, pad=1, relu=True, device="cuda", engine="tc16")
    x = E.Act(torch.randn(B, H, W, Cin, generator=g).cuda(), 0, None, torch.zeros(2, B, H, W, Cin, device="cuda", dtype=torch.float16))
    E.split_lo(x)
    out = E.Act(torch.zeros(B, H, W, Cout, device="cuda"), 0, None, torch.zeros(2, B, H, W, Cout, device="cuda", dtype=torch.float16))
    for cfg in os.environ.get("TRACE_CFGS", "0:0,1:128,2:128").split(","):
        cg, bn = (int(v) for v in cfg.split(":"))
        os.environ["VD3D_TC_CG"] = str(cg)
        layer.bn_tile = bn
        layer(x, out); torch.cuda.synchronize()
        tr = torch.zeros(5, N, dtype=torch.int64, device="cuda")
        lib.vd3d_tc_set_trace(tr.data_ptr(), N)
        layer(x, out); torch.cuda.synchronize()
        lib.vd3d_tc_set_trace(None, 0)
        t = tr.cpu().numpy().astype(np.float64)
        n = int((t[0] > 0).sum())
        t = t[:, :n]
        t0 = t[0, 0]
        lat = t[3] - t[0]                  # stage free -> its MMAs start
        wait = t[2] - t[3]                 # issue start -> look-ahead wait for the next stage starts (= first half of the MMAs issued)
        issue = t[4] - t[3]                # 12 MMAs + look-ahead (wait, fence, descriptors) + commit
        per = np.diff(t[4])
        sl = slice(20, min(n, 400))
        print(f"{name:30s} cg={cg} bn={bn:3d} n={n}  period {np.median(per[sl]):7.0f}  free->mma-start {np.median(lat[sl]):7.0f}  "
              f"issue-loads {np.median((t[1]-t[0])[sl]):5.0f}  half-issue {np.median(wait[sl]):7.0f}  mma-issue {np.median(issue[sl]):6.0f} clk", flush=True)
        k = 40
        print("   k-block:", " ".join(f"{int(v):6d}" for v in range(k, k + 8)))
        for r, lab in enumerate(["free", "issued", "lookahd", "mmastart", "mmadone"]):
            print(f"   {lab:8s}", " ".join(f"{int(v - t0):6d}" for v in t[r, k:k + 8]))
