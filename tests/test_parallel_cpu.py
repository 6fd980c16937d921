"""N > 1 path on CPU (gloo, world_size 2): sharding + the single all-gather of fixed-capacity detection records
(visualdet3d_b200/parallel.py).  Parity target = concatenation of the per-image results in global batch order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from visualdet3d_b200 import parallel


def _fake_results(n_images, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n_images):
        k = int(torch.randint(0, 7, (1,), generator=g))
        out.append((torch.rand(k, generator=g), torch.randn(k, 11, generator=g) * 100, torch.randint(0, 3, (k,), generator=g)))
    return out


def test_pack_unpack_roundtrip_and_capacity():
    res = _fake_results(5, 0) + [(torch.zeros(0), torch.zeros(0, 11), torch.zeros(0, dtype=torch.int64))]
    buf = parallel.pack_records(res, 8, "cpu")
    assert buf.shape == (6, 1 + 8 * 13)
    back = parallel.unpack_records(buf)
    for (s, b, c), (s2, b2, c2) in zip(res, back):
        assert torch.equal(s, s2) and torch.equal(b, b2) and torch.equal(c, c2) and c2.dtype == torch.int64
    assert back[-1][0].shape == (0,) and back[-1][1].shape == (0, 11)
    with pytest.raises(RuntimeError):
        parallel.pack_records([(torch.rand(9), torch.rand(9, 11), torch.zeros(9, dtype=torch.int64))], 8, "cpu")


def test_shard_range_covers_batch():
    for n, w in [(64, 8), (10, 4), (3, 8), (8, 1)]:
        spans = [parallel.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        allres = _fake_results(8, 123)                      # the "global batch" every rank can recompute
        lo, hi = parallel.shard_range(8, rank, world)
        got = parallel.all_gather_detections(allres[lo:hi], 8, "cpu")
        ok = len(got) == 8 and all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
                                   for a, b in zip(allres, got))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_all_gather_detections_gloo_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in ps]
    [p.join(60) for p in ps]
    assert sorted(res) == [(0, True), (1, True)]


def test_geometry_records_and_guard_flags():
    """The 20-column record form (post-forward columns appended by vd3d_pack_records_geo) unpacks into the plain triples plus
    (box3d, theta, box2d); count == -1 (capacity) and -2 (fp16-range guard of the tensor-core engine) raise instead of returning rows."""
    from visualdet3d_b200._lib import Vd3dError
    g = torch.Generator().manual_seed(3)
    kmax, R = 4, parallel.REC_GEO
    buf = torch.zeros(2, 1 + kmax * R)
    rows = torch.randn(3, R, generator=g)
    rows[:, 12] = torch.tensor([0.0, 2.0, 1.0])
    buf[0, 0] = 3
    buf[0, 1:1 + 3 * R] = rows.reshape(-1)
    geo = []
    out = parallel.unpack_records(buf, R, geo)
    assert len(out) == 2 and len(geo) == 2 and out[1][0].shape == (0,) and geo[1][0].shape == (0, 7)
    s, b, c = out[0]
    assert torch.equal(s, rows[:, 11]) and torch.equal(b, rows[:, :11]) and c.tolist() == [0, 2, 1]
    box3d, theta, box2d = geo[0]
    assert torch.equal(box3d[:, :2], rows[:, 13:15]) and torch.equal(box3d[:, 2:], rows[:, 6:11])
    assert torch.equal(theta, rows[:, 15]) and torch.equal(box2d, rows[:, 16:20])
    assert [tuple(t.shape) for t in parallel.unpack_records(buf, R)[0]] == [(3,), (3, 11), (3,)]     # geometry list optional
    buf[1, 0] = -1
    with pytest.raises(RuntimeError):
        parallel.unpack_records(buf, R)
    buf[1, 0] = -2
    with pytest.raises(Vd3dError, match="fp16 range"):
        parallel.unpack_records(buf, R)
