"""-m gpu, needs >= 2 GPUs (skipped otherwise; run with `gpurun --gpus 2`): the N > 1 path on real hardware over NCCL
(tests/workers/nccl_worker.py under torch.distributed.run, 2 ranks).  The CPU counterpart is tests/test_parallel_cpu.py (gloo)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_nccl_gather_and_pipeline():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "workers", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("NCCL_JSON ")]
    assert r.returncode == 0 and lines, f"rc {r.returncode}\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    out = json.loads(lines[-1][len("NCCL_JSON "):])
    print(out)
    assert out["world"] == 2
    assert out["own_slice_bit_exact"] and out["other_ranks_recomputed_bit_exact"] and out["unpacked_equals_local"]
    assert out["pipeline_post_opt_geometry_ok"]
    assert out["stereo_detections"] > 0 and out["mono_detections"] > 0
