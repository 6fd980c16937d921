"""Per-layer-group error budget of a 2-pass product (SURVEY / VERDICT r1 item 9): what does dropping the A_lo * W_hi MMA (activations seen
with 11 significant bits) cost in final box / score error, and what does it save?  Stereo3D, BASELINE configs[1] shape, batch 8.
usage: VD3D_PLANES=0 python tools/error_budget.py  -> one line per layer group (max |d score|, max |d box| over rows kept by both, set changes, ms)"""
import os, sys
os.environ["VD3D_PLANES"] = "0"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from visualdet3d_b200 import synth, engine as E
from visualdet3d_b200.detectors import build_synthetic_stereo3d

det = build_synthetic_stereo3d(seed=0)[0].cuda().eval()
B, H, W = 8, 384, 1280
left, right, P2, _ = synth.synth_stereo_inputs(B, H, W, seed=11)
l, r, p = left.cuda(), right.cuda(), P2.cuda()
pl = det.prepare()


def convs(obj):
    out = []
    if isinstance(obj, E.ConvLayer):
        out.append(obj)
    elif isinstance(obj, dict):
        for v in obj.values():
            out += convs(v)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            out += convs(v)
    elif hasattr(obj, "__dict__"):
        for v in vars(obj).values():
            if isinstance(v, (E.ConvLayer, dict, list, tuple)) or type(v).__name__ in ("GhostRunner", "ResNetRunner"):
                out += convs(v)
    return out


bb = pl["backbone"]
groups = {
    "backbone stem..layer1 (64 ch)": convs(bb.stages[0]),
    "backbone layer2 (128 ch)": convs(bb.stages[1]),
    "backbone layer3 (256 ch)": convs(bb.stages[2]),
    "neck ghost / basic blocks (<= 384 ch)": convs([pl["g4"], pl["bb4"], pl["g8"], pl["bb8"], pl["g16"]]),
    "neck 1152 -> 1152 pair": convs(pl["bb16"]),
    "cls tower": convs(pl["cls"]),
    "reg tower 1408 -> 1408 x3": convs([pl["reg0"], pl["reg_bb"]]),
    "reg output conv": convs(pl["reg_out"]),
}


def run():
    with torch.no_grad():
        for _ in range(2):
            det.launch(l, r, p)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            dec = det.launch(l, r, p)
        b.record()
        torch.cuda.synchronize()
        res = dec.results()
        anchors = [dec.anchor[i, :len(res[i][0])].clone() for i in range(B)]
    return [(s.clone(), bx.clone(), an) for (s, bx, c), an in zip(res, anchors)], a.elapsed_time(b) / 5


ref, ms3 = run()
print(f"3 passes everywhere: {ms3:.3f} ms / step, {sum(len(x[0]) for x in ref)} detections")
all_tc = [c for g in groups.values() for c in g if c.engine == "tc16"]
for name, layers in list(groups.items()) + [("ALL tensor-core convs", all_tc)]:
    layers = [c for c in layers if c.engine == "tc16"]
    for c in layers:
        c.passes = 2
    got, ms = run()
    for c in layers:
        c.passes = 3
    ds = db = 0.0
    lost = gained = 0
    for (s, bx, an), (rs, rb, ran) in zip(got, ref):
        pos = {int(a): i for i, a in enumerate(ran.tolist())}
        both = [(i, pos[int(a)]) for i, a in enumerate(an.tolist()) if int(a) in pos]
        gained += len(an) - len(both)
        lost += len(ran) - len(both)
        if both:
            i0 = torch.tensor([x for x, _ in both], device=s.device)
            i1 = torch.tensor([y for _, y in both], device=s.device)
            ds = max(ds, float((s[i0] - rs[i1]).abs().max()))
            db = max(db, float((bx[i0] - rb[i1]).abs().max()))
    print(f"2 passes in {name:40s} ({len(layers):2d} convs): max |d score| {ds:.2e}  max |d box| {db:.2e}  kept-set changes -{lost} +{gained}  {ms:.3f} ms ({100 * (ms - ms3) / ms3:+.1f} %)", flush=True)
