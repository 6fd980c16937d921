#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 visualDet3D hot path (contract: see the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--batch 8]
    torchrun --nproc-per-node N bench.py --gpus N ...        (one rank per GPU, NCCL)

A "step" = one YOLOStereo3D forward (backbone -> cost volumes -> neck -> head -> decode -> NMS) over a batch of
`--batch` synthetic 384x1280 stereo pairs per GPU (BASELINE.json configs[1]); ranks hold disjoint pairs (weak scaling)
and exchange only one all-gather of detection records.  Prints ONE JSON line on rank 0.

  value     : whole-job stereo pairs/s, inputs resident in HBM, timed with CUDA events, max over ranks
  e2e       : same metric through the public detector API with HOST (pinned) inputs: H2D copies, forward, D2H of results
  roofline  : scale-4 PSMCosine kernel (dominant cost-volume kernel): algorithmic bytes / CUDA-event time vs measured HBM peak
  cpu_baseline : the CPU oracle port (oracle/torch_port.py, the reference's algorithm in fp32 PyTorch ops) on this host's cores
  --impl reference : times that CPU implementation alone (the reference itself is Python and cannot travel to the GPU box)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "synthetic_384x1280_stereo_pairs_per_sec"
UNIT = "pairs/s"
H, W = 384, 1280
PSM4_BYTES_PER_PAIR = 4 * (H // 4) * (W // 4) * (2 * 64 + 24)   # 18,677,760 B
PSM4_NCU_TRAFFIC_B8 = 137_400_000                                # DRAM bytes per launch of the scale-4 PSMCosine kernel at B = 8 (ncu)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock / power / throttle-reason sampling DURING the timed region (B200_PROFILING.md): NVML polled every 5 ms from a
    thread (nvidia_ml_py), falling back to `nvidia-smi -lms` when NVML is not importable."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.h, self.stop_flag, self.samples = None, None, False, []
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
                pw = n.nvmlDeviceGetPowerUsage(self.h) / 1e3
                try:
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((sm, pw, rs))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.nvml is not None:
            self.stop_flag, self.samples = False, []
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            n = self.nvml
            try:
                mx = float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM))
            except Exception:
                mx = None
            sm = [float(a) for a, _, _ in self.samples]
            pw = [b for _, b, _ in self.samples]
            bits = 0
            for _, _, r in self.samples:
                bits |= int(r)
            reasons = sorted(k for k, v in self.BITS.items() if bits & v)
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None, "sm_max_mhz": mx,
                    "power_w": statistics.median(pw) if pw else None, "reasons": reasons, "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def pick_cpu_threads():
    """Thread count for the CPU arm: the candidate (all usable cores, 64, 32, 16, 8) that runs a small conv stack fastest.
    On the GPU boxes `torch.set_num_threads(nproc=128)` is pathologically slow for oneDNN convs (28 s per pair), so the
    count is calibrated rather than assumed; the chosen value is what `cores` reports."""
    import torch
    import torch.nn.functional as F
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    x = torch.randn(2, 64, 96, 320)
    w = torch.randn(64, 64, 3, 3)
    best, best_t = avail, None
    for n in sorted({avail, 64, 32, 16, 8}):
        if n > avail:
            continue
        torch.set_num_threads(n)
        F.conv2d(x, w, padding=1)
        t0 = time.perf_counter()
        for _ in range(4):
            F.conv2d(x, w, padding=1)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def cpu_forward_rate(pairs: int, iters: int, threads: int):
    """Oracle port on the host cores: `iters` forwards of `pairs` pairs at 384x1280 -> (pairs/s, seconds per forward)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch_port as tp
    from visualdet3d_b200 import synth
    from visualdet3d_b200.detectors import build_synthetic_stereo3d
    torch.set_num_threads(threads)
    det, sd, cfg, (pm, ps) = build_synthetic_stereo3d(seed=0)
    left, right, P2, P3 = synth.synth_stereo_inputs(pairs, H, W, seed=1)
    tp.stereo3d_forward(sd, left, right, P2, cfg, pm, ps)          # warm-up (oneDNN primitive creation)
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        tp.stereo3d_forward(sd, left, right, P2, cfg, pm, ps)
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return pairs / med, med


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = pick_cpu_threads()
    pairs = 1
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch_port as tp
    from visualdet3d_b200 import synth
    from visualdet3d_b200.detectors import build_synthetic_stereo3d
    torch.set_num_threads(cores)
    det, sd, cfg, (pm, ps) = build_synthetic_stereo3d(seed=0)
    left, right, P2, P3 = synth.synth_stereo_inputs(pairs, H, W, seed=1)
    for _ in range(max(1, min(args.warmup, 2))):
        tp.stereo3d_forward(sd, left, right, P2, cfg, pm, ps)
    steps = max(1, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(steps):
        tp.stereo3d_forward(sd, left, right, P2, cfg, pm, ps)
    dt = time.perf_counter() - t0
    v = pairs * steps / dt
    sample = f"{steps} forwards of {pairs} pair(s) 384x1280 (bounded sample of the batch-{args.batch} workload)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": f"YOLOStereo3D forward, stereo 384x1280, ResNet-34 (CPU, {sample})"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="stereo pairs per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-mode", action="store_true", help="device-resident steps only (for ncu): no e2e leg, no CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from visualdet3d_b200 import _lib, synth, parallel
    from visualdet3d_b200.detectors import build_synthetic_stereo3d

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args.warmup = max(args.warmup, 3)
    B = args.batch
    kmax = 512

    det, sd, cfg, _ = build_synthetic_stereo3d(seed=0)
    det = det.to(dev).eval()
    # every rank owns its own B pairs of the global batch (weak scaling): different seeds per rank
    left, right, P2, P3 = synth.synth_stereo_inputs(B, H, W, seed=1 + rank)
    hl, hr, hp = left.pin_memory(), right.pin_memory(), P2.pin_memory()
    dl, dr, dp = hl.to(dev), hr.to(dev), hp.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        dec = det.launch(dl, dr, dp)
        return dec

    from visualdet3d_b200.pipeline import StreamedInference
    pipe = StreamedInference(det, B, H, W, kmax=kmax, world=world)

    def run_e2e(nsteps):
        """`nsteps` batches through the public host-fed pipeline: every batch pays its pinned-host -> device copy and the
        device -> host read of the gathered detection records; copy of batch i+1 overlaps the forward of batch i."""
        out = None
        prev = None
        for _ in range(nsteps):
            t = pipe.submit(hl, hr, hp)
            if prev is not None:
                out = pipe.collect(prev)
            prev = t
        out = pipe.collect(prev)
        return out[rank * B:(rank + 1) * B], out

    with torch.no_grad():
        for _ in range(args.warmup):
            step_device()
            if not args.profile_mode:
                res, _h = run_e2e(1)
        if args.profile_mode:
            torch.cuda.synchronize()
            _lib.launch_count_reset()
            for _ in range(args.steps):
                step_device()
            torch.cuda.synchronize()
            print(json.dumps({"profile_mode": True, "launches_per_step": _lib.launch_count() / args.steps}))
            return
        # ---------------- device-resident timing ----------------------------------------------------------------
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        det.profile_events = []
        _lib.launch_count_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            dec = step_device()
            # the all-gather belongs to the step: device-built record block -> NCCL, no host synchronisation
            parallel.all_gather_records(parallel.pack_records_device(dec, kmax))
        e1.record()
        barrier()
        launches = _lib.launch_count()
        ms_dev = e0.elapsed_time(e1)
        psm_ms = [a.elapsed_time(b) for a, b in det.profile_events]
        det.profile_events = None
        clocks = sampler.stop() if rank == 0 else None
        # ---------------- end-to-end timing (host inputs) -----------------------------------------------------------
        barrier()
        t0 = time.perf_counter()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        res, host = run_e2e(args.steps)
        e3.record()
        barrier()
        ms_e2e = max(e2.elapsed_time(e3), 1e3 * (time.perf_counter() - t0))      # device time and host wall clock: the larger one
    t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pairs_total = B * world * args.steps
    value = pairs_total / (ms_dev / 1e3)
    e2e_v = pairs_total / (ms_e2e / 1e3)
    peak, peak_kind = measured_peaks()
    psm_avg_ms = statistics.mean(psm_ms) if psm_ms else None
    achieved = (PSM4_BYTES_PER_PAIR * B / 1e9) / (psm_avg_ms / 1e3) if psm_avg_ms else None
    ndet = sum(len(r[0]) for r in res)
    h2d, d2h = int(pipe.h2d_bytes), int(pipe.d2h_bytes)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"YOLOStereo3D forward, batch {B} stereo 384x1280 per GPU, ResNet-34, random-init seeded weights",
                   "global_batch": B * world, "parallelism": f"dp{world}", "l2": "inputs+weights (524 MB/step) exceed the 126 MB L2; no explicit flush",
                   "conv_engine": os.environ.get("VD3D_CONV_ENGINE", "default"), "detections_per_step": ndet},
        "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                "api": "visualdet3d_b200.pipeline.StreamedInference (pinned host batches, double-buffered H2D on a copy stream, async D2H of the gathered records)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": ("psm_cosine_tc_kernel (scale-4 PSMCosine, tcgen05 on fp16 hi/lo planes)"
                                if (os.environ.get("VD3D_PSM_ENGINE", "tc") == "tc" and os.environ.get("VD3D_CONV_ENGINE", "tc16") == "tc16")
                                else "psm_cosine_nhwc_v4_kernel<64> (scale-4 PSMCosine, SIMT)"), "bound": "hbm", "achieved": achieved, "peak": peak,
                     "peak_kind": peak_kind, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                     "avg_launch_ms": psm_avg_ms, "algorithmic_bytes_per_launch": PSM4_BYTES_PER_PAIR * B,
                     # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at B = 8 from the committed `ncu --set full` capture
                     # (profiles/r01_ncu_psm_cosine_tc.txt: 125.9 MB read + 11.5 MB written, 27.6 us cold); null for other batch sizes
                     "traffic": (PSM4_NCU_TRAFFIC_B8 if (B == 8 and os.environ.get("VD3D_PSM_ENGINE", "tc") == "tc") else None),
                     "traffic_source": "ncu --set full, one launch, profiles/r01_ncu_psm_cosine_tc.txt"},
    }
    if not args.no_cpu_baseline and world == 1:          # the CPU arm is timed on rank 0 at N = 1 only (the driver runs --impl reference for every N)
        cores = pick_cpu_threads()
        v, sec = cpu_forward_rate(1, 3, cores)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"3 forwards of 1 pair 384x1280 (median {sec:.2f} s), oracle/torch_port.py on {cores} threads"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
