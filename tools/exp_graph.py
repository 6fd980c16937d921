"""Eager launches vs one CUDA graph of the whole forward (backbone .. NMS .. record block): how much of the step is launch gaps?

    python tools/exp_graph.py [--config stereo|gac|monoflex|km3d|yolo3d] [--steps 20]

Prints one JSON line per config: ms per step eager / graph replay, and whether the replayed record block equals the eager one bit for bit.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from visualdet3d_b200 import parallel, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="stereo")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=None)
    args = ap.parse_args()
    metric, unit, H, W, defB, text = bench.CONFIGS[args.config]
    B = args.batch or defB
    dev = torch.device("cuda", 0)
    det = bench.build_detector(args.config).to(dev).eval()
    if args.config == "stereo":
        l, r, p2, _ = synth.synth_stereo_inputs(B, H, W, seed=1)
        imgs = [l.to(dev), r.to(dev)]
    else:
        im, p2 = synth.synth_mono_inputs(B, H, W, seed=1)
        imgs = [im.to(dev)]
    p2 = p2.to(dev)
    rec = torch.empty(B, 1 + 512 * parallel.REC, device=dev)

    def step():
        dec = det.launch(*imgs, p2)
        parallel.pack_records_device(dec, 512, out=rec)

    def timed(fn, n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    with torch.no_grad():
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        want = rec.clone()
        ms_eager = timed(step, args.steps)
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            step()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            step()
        rec.zero_()
        g.replay()
        torch.cuda.synchronize()
        same = bool(torch.equal(rec, want))
        ms_graph = timed(g.replay, args.steps)
        ms_eager2 = timed(step, args.steps)
    print(json.dumps({"config": args.config, "batch": B, "ms_eager": ms_eager, "ms_graph": ms_graph, "ms_eager_again": ms_eager2,
                      "graph_equals_eager": same, "gain": ms_eager / ms_graph}))


if __name__ == "__main__":
    main()
