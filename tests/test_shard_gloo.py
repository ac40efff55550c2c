"""CPU, world_size 2 over gloo: the host-side sharding / gather logic of the multi-GPU batch path.
(The fold itself needs a GPU; a stand-in fold function is injected here -- the product has no CPU fold.)"""
import os
import socket

import pytest
import torch.multiprocessing as mp

from ccj_b200.shard import LocalCounter, StoreCounter, fold_dealt, fold_sharded, partition


def test_partition_covers_and_balances():
    lens = [150] * 64
    for w in (1, 2, 4, 8):
        b = partition(lens, w)
        assert b[0][0] == 0 and b[-1][1] == 64
        assert all(b[r][1] == b[r + 1][0] for r in range(w - 1))
        assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1
    ragged = [30, 200, 40, 50, 180, 20, 10, 160]
    b = partition(ragged, 2)
    cost = lambda s, e: sum(x ** 5 for x in ragged[s:e])
    assert abs(cost(*b[0]) - cost(*b[1])) < 0.6 * sum(x ** 5 for x in ragged)
    assert partition([10], 4)[-1][1] == 1
    assert partition([], 2) == [(0, 0), (0, 0)]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seqs = ["ACGU" * (3 + x % 5) for x in range(11)]
    fake_fold = lambda batch: [(s, len(s), rank) for s in batch]
    out = fold_sharded(seqs, fake_fold, rank, world)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, out))


def test_two_ranks_gather_in_input_order():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seqs = ["ACGU" * (3 + x % 5) for x in range(11)]
    assert res[0] == res[1]
    assert [x[0] for x in res[0]] == seqs
    assert {x[2] for x in res[0]} == {0, 1}


def test_dealt_single_process_covers_every_index_once():
    c = LocalCounter()
    out = fold_dealt(37, 8, c, lambda lo, hi: [x * x for x in range(lo, hi)])
    assert out == [(x, x * x) for x in range(37)]
    assert fold_dealt(37, 8, c, lambda lo, hi: [0] * (hi - lo)) == []   # counter exhausted


def _deal_worker(rank, world, port, q):
    import time
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    counter = StoreCounter(dist.distributed_c10d._get_default_store(), "ccj_test_next")

    def fake_fold(lo, hi):   # rank 1 is the "power-capped" GPU: it must end up with fewer chunks
        time.sleep(0.05 if rank == 1 else 0.005)
        return [(x, rank) for x in range(lo, hi)]
    dist.barrier()
    mine = fold_dealt(203, 7, counter, fake_fold)
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, parts))


def test_two_ranks_deal_dynamically():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_deal_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == res[1]
    flat = sorted(x for part in res[0] for x in part)
    assert [x[0] for x in flat] == list(range(203))            # every index exactly once
    assert all(x[1][0] == x[0] for x in flat)
    n0, n1 = len(res[0][0]), len(res[0][1])
    assert n0 + n1 == 203 and n0 > n1 > 0                      # the slower rank drew fewer chunks
