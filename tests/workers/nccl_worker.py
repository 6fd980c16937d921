"""Worker of tests/test_nccl_gpu.py: launched by torch.distributed.run with 2 ranks on 2 GPUs of one box (NCCL).

Checks on real hardware what the gloo test can only check with fake tensors:
  1. the ONE collective of the path: every rank's slice of the all-gathered record block equals its local `pack_records_device` block
     bit for bit, and rank 0 recomputes rank 1's batch itself and finds the same bits in rank 1's slice;
  2. BASELINE configs[2] with `post_optimization=True` (the device hill-climbing kernel) through the host-fed pipeline at world = 2
     with the post-forward geometry columns: the global result list == the per-rank `forward_batch` results in rank order.
Rank 0 prints one `NCCL_JSON {...}` line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from visualdet3d_b200 import parallel, synth
    from visualdet3d_b200.detectors import build_synthetic_mono3d, build_synthetic_stereo3d
    from visualdet3d_b200.pipeline import StreamedInference
    out = {"world": world}
    # ---- 1. stereo: raw collective ------------------------------------------------------------------------------------
    det = build_synthetic_stereo3d(seed=0)[0].to(dev).eval()
    B, H, W, kmax = 2, 96, 320, 64
    inputs = lambda r: synth.synth_stereo_inputs(B, H, W, seed=40 + r)
    left, right, P2, _ = inputs(rank)
    with torch.no_grad():
        dec = det.launch(left.to(dev), right.to(dev), P2.to(dev))
        rec = parallel.pack_records_device(dec, kmax)
        local_res = [(s.cpu().clone(), b.cpu().clone(), c.cpu().clone()) for s, b, c in dec.results()]
        gathered = parallel.all_gather_records(rec)
        torch.cuda.synchronize()
        ok_own = torch.equal(gathered[rank * B:(rank + 1) * B], rec)
        ok_other = True
        if rank == 0:
            for r in range(1, world):
                l2, r2, p2, _ = inputs(r)
                rec_r = parallel.pack_records_device(det.launch(l2.to(dev), r2.to(dev), p2.to(dev)), kmax)
                torch.cuda.synchronize()
                ok_other = ok_other and torch.equal(gathered[r * B:(r + 1) * B], rec_r)
    glob = parallel.unpack_records(gathered.cpu())
    ok_unpack = all(torch.equal(a, b) for x, y in zip(glob[rank * B:(rank + 1) * B], local_res) for a, b in zip(x, y))
    n_det = sum(len(g[0]) for g in glob)
    # ---- 2. mono + post_optimization + geometry through the pipeline at world 2 -------------------------------------------
    mdet = build_synthetic_mono3d("Yolo3D", seed=0)[0].to(dev).eval()
    mdet.post_optimization = True
    img, mP2 = synth.synth_mono_inputs(B, H, W, seed=60 + rank)
    pipe = StreamedInference(mdet, B, H, W, kmax=kmax, world=world, geometry=True)
    tickets = [pipe.submit(img.pin_memory(), mP2.pin_memory(), original_P=mP2.pin_memory()) for _ in range(2)]
    got = [pipe.collect(t) for t in tickets][-1]
    geo = pipe.last_geometry
    with torch.no_grad():
        ref = mdet.forward_batch(img.to(dev), mP2.to(dev))
    ok_pipe = len(got) == world * B and len(geo) == world * B and all(
        torch.equal(a, b.cpu()) for x, y in zip(got[rank * B:(rank + 1) * B], ref) for a, b in zip(x, y))
    n_mono = sum(len(g[0]) for g in got)
    flags = torch.tensor([int(ok_own), int(ok_other), int(ok_unpack), int(ok_pipe)], device=dev, dtype=torch.int32)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.update(own_slice_bit_exact=bool(flags[0]), other_ranks_recomputed_bit_exact=bool(flags[1]), unpacked_equals_local=bool(flags[2]),
                   pipeline_post_opt_geometry_ok=bool(flags[3]), stereo_detections=n_det, mono_detections=n_mono)
        print("NCCL_JSON " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
