"""Where does the time of the fused deformable conv go?  One DLA up-sampling layer shape, CUDA-event timing under knock-outs (results wrong).
usage: python tools/exp_dcn.py [C] [H] [W] [reps]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visualdet3d_b200 import engine as E

C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
H = int(sys.argv[2]) if len(sys.argv) > 2 else 96
W = int(sys.argv[3]) if len(sys.argv) > 3 else 320
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
B, Co = 8, C
g = torch.Generator().manual_seed(0)
layer = E.DeformConvLayer(torch.randn(Co, C, 3, 3, generator=g) / np.sqrt(9 * C), torch.randn(Co, generator=g), torch.randn(27, C, 3, 3, generator=g) * 0.02,
                          torch.randn(27, generator=g) * 0.3, None, stride=1, pad=1, dil=1, relu=True, device="cuda")
planes = lambda c: torch.zeros(2, B, H, W, c, device="cuda", dtype=torch.float16)
x = E.split_lo(E.Act(torch.randn(B, H, W, C, generator=g).cuda(), 0, None, planes(C)))
out = E.Act(torch.zeros(B, H, W, Co, device="cuda"), 0, None, planes(Co))
ar = E.Arena()
flush = torch.empty(64 * 1024 * 1024, device="cuda")


def run(label, env):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        ts, to = [], []
        for i in range(reps + 2):
            flush.zero_()
            a, m, b = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            om = layer.off_conv(x, ar.act("om", (B, H, W, layer.n_off_pad), "cuda"))
            m.record()
            layer(x, out, ar, "t")
            b.record()
            torch.cuda.synchronize()
            if i >= 2:
                to.append(a.elapsed_time(m) * 1e3)
                ts.append(m.elapsed_time(b) * 1e3)
        print(f"C={C} {H}x{W}  {label:62s} layer (offset conv + main) {np.median(ts):8.1f} us   offset conv alone {np.median(to):7.1f} us", flush=True)
    except Exception as e:
        print(f"{label}: FAILED {e!r}"[:200], flush=True)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


run("staged fused kernel (default)", {})
run("unfused: column planes + 1x1 conv", {"VD3D_DCN_FUSED": 0})
if C == 64:
    run("fused, corners gathered from global memory", {"VD3D_DCN_STAGED": 0})
for bits, lab in ((1, "knock-out: no corner loads"), (2, "knock-out: no offset / mask loads"), (4, "knock-out: no operand stores"), (8, "knock-out: one MMA per k-block"),
                  (3, "knock-out: no corner loads, no offset loads"), (7, "knock-out: no corner / offset loads, no operand stores"),
                  (15, "knock-out: all four")):
    run(lab, {"VD3D_DF_DEBUG": bits})
run("knock-out: no epilogue output", {"VD3D_TC_DEBUG": 16})
run("knock-out: all four + no epilogue output", {"VD3D_DF_DEBUG": 15, "VD3D_TC_DEBUG": 16})
