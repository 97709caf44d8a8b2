"""world_size-2 gloo test (CPU) of the multi-GPU host logic: round-robin image sharding and the
result-record all_gather that is the path's only exchange."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from densecap_b200 import parallel as PL

B, P, T = 2, 5, 3


def _fake_results(rank):
    g = torch.Generator().manual_seed(100 + rank)
    nf = torch.tensor([3 + rank, 1], dtype=torch.int32)
    boxes = torch.rand((B * P, 4), generator=g) * 700
    scores = torch.randn((B * P,), generator=g)
    seq = torch.randint(1, 10499, (B * P, T), generator=g, dtype=torch.int64)
    return nf, boxes, scores, seq


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = PL.shard_indices(7, rank, world)
        rec = PL.pack_result_record(*_fake_results(rank))
        assert rec.numel() == PL.record_numel(B, P, T)
        allrec = PL.gather_records(dist, rec)
        ok = True
        for r in range(world):
            nf, boxes, scores, seq = PL.unpack_result_record(allrec[r], B, P, T)
            enf, eb, es, eq = _fake_results(r)
            ok &= bool(torch.equal(nf, enf) and torch.equal(boxes, eb) and torch.equal(scores, es) and torch.equal(seq, eq))
        q.put((rank, mine, ok))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_and_gather_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    assert all(r[2] for r in res)
    # every image is owned exactly once
    owned = sorted(res[0][1] + res[1][1])
    assert owned == list(range(7))


def test_record_round_trip_single_process():
    nf, boxes, scores, seq = _fake_results(0)
    rec = PL.pack_result_record(nf, boxes, scores, seq)
    a, b, c, d = PL.unpack_result_record(rec, B, P, T)
    assert torch.equal(a, nf) and torch.equal(b, boxes) and torch.equal(c, scores) and torch.equal(d, seq)
