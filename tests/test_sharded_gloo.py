"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: image sharding and the final detections all-gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yolov4_b200.sharded import allgather_detections, shard_range


def test_shard_range_partitions_exactly():
    for n in (1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_shard(rank, b_local, cap):
    g = torch.Generator().manual_seed(100 + rank)
    counts = torch.randint(0, cap, (b_local,), generator=g, dtype=torch.int32)
    counts[rank % b_local] = 0                     # an image without detections -> None
    rows = torch.rand((b_local, cap, 7), generator=g)
    return rows, counts


def _worker(rank, world, port, b_local, cap, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows, counts = _fake_shard(rank, b_local, cap)
    out = allgather_detections(rows, counts)
    ok = len(out) == world * b_local
    for r in range(world):
        rr, cc = _fake_shard(r, b_local, cap)
        for b in range(b_local):
            o = out[r * b_local + b]
            k = int(cc[b])
            ok = ok and ((o is None) if k == 0 else (o is not None and torch.equal(o, rr[b, :k])))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_allgather_detections_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 3, 17, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
