"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the sharding plan and the random-count exchange."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dwgsim_b200 import shard  # noqa: E402


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ex = shard.make_exchange()
        mine = shard.plan(total, batch, rank, world)
        # deterministic fake "random pairs in batch b"
        fake = lambda b: (b * 7919 + 13) % 101
        base, log = 0, []
        for rnd in range(shard.n_rounds(total, batch, world)):
            own = [m for m in mine if m[0] == rnd]
            cnt = fake(own[0][1] // batch) if own else 0
            before, tot = ex(rnd, cnt)
            if own:
                log.append((own[0][1], own[0][2], base + before))
            base += tot
        q.put((rank, log, base))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total,batch", [(10_000, 1024), (4096, 1024), (1000, 4096), (12_345, 100)])
def test_plan_and_exchange_world2(total, batch):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, total, batch, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    # serial truth
    n_batches = (total + batch - 1) // batch
    fake = lambda b: (b * 7919 + 13) % 101
    serial, base = {}, 0
    for b in range(n_batches):
        serial[b * batch] = base
        base += fake(b)
    seen = {}
    for rank, log, final in res:
        assert final == base
        for first, n, rb in log:
            assert first not in seen
            seen[first] = (n, rb)
    assert sorted(seen) == sorted(serial)
    assert sum(n for n, _ in seen.values()) == total
    for first, (n, rb) in seen.items():
        assert rb == serial[first], (first, rb, serial[first])


def test_interleave():
    assert shard.interleave([[b"a", b"c", b"e"], [b"b", b"d"]]) == b"abcde"
    assert shard.interleave([[b"a"], []]) == b"a"
    assert shard.interleave([[b"a", b"c"], [b"b", b"d"]]) == b"abcd"
