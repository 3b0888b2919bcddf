"""world_size-2 gloo test of the detection all-gather (the only collective of the path) and the batch
sharding rules — runs on CPU."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import yolox_b200 as yb
from yolox_b200 import dist as ydist


def test_shard_range_covers_batch():
    for n in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [ydist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_roundtrip():
    det = torch.randn(5, 300, 7)
    cnt = torch.tensor([0, 1, 300, 17, 256], dtype=torch.int32)
    d2, c2 = ydist.unpack_detections(ydist.pack_detections(det, cnt), 300)
    assert torch.equal(d2, det) and torch.equal(c2, cnt)


def test_wrap_i32_matches_device_counter():
    wrap_i32 = ydist.wrap_i32
    assert wrap_i32(5) == 5 and wrap_i32(2 ** 31 - 1) == 2 ** 31 - 1
    assert wrap_i32(2 ** 31) == -2 ** 31 and wrap_i32(2 ** 32 + 7) == 7


def _worker(rank, world, port, n_images, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = -(-n_images // world)
    s, e = ydist.shard_range(n_images, rank, world)
    g = torch.Generator().manual_seed(100)
    det_global = torch.randn(n_images, 300, 7, generator=g)
    cnt_global = torch.randint(0, 301, (n_images,), generator=g).to(torch.int32)
    det, cnt = det_global[s:e], cnt_global[s:e]
    if e - s < per:  # padding rows
        det = torch.cat([det, det[-1:].expand(per - (e - s), -1, -1)], 0)
        cnt = torch.cat([cnt, cnt[-1:].expand(per - (e - s))], 0)
    det_all, cnt_all = ydist.all_gather_detections(det.contiguous(), cnt.contiguous())
    det_all, cnt_all = ydist.gathered_for_images(det_all, cnt_all, n_images, world, per)
    ok = torch.equal(det_all, det_global) and torch.equal(cnt_all, cnt_global)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_all_gather_detections_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    for n_images in (8, 7):
        out = ctx.Manager().dict()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, out)) for r in range(2)]
        [p.start() for p in procs]
        [p.join(120) for p in procs]
        assert all(p.exitcode == 0 for p in procs)
        assert out[0] and out[1]
        port += 1


def _records_worker(rank, world, port, out):
    from yolox_b200.evaluator import gather_records
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = [dict(image_id=100 * rank + i, category_id=3 + i, bbox=[1.5 * i, 2.0, 3.25, 4.0 + rank], score=0.125 * (i + 1),
                 segmentation=[]) for i in range(3 if rank == 0 else 0)]      # rank 1 has nothing to report
    got = gather_records(mine, dst=0)
    if rank == 0:
        ok = len(got) == world and got[0] == mine and got[1] == []
    else:
        ok = got == []
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_evaluator_gather_records_world2():
    """COCOEvaluator's distributed collect (coco_evaluator.py:127): ragged per-rank record lists arrive on rank 0."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_records_worker, args=(r, 2, port, out)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert out[0] and out[1]
