"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: clip sharding + max-over-ranks timing reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vnext_b200.dist import max_over_ranks, shard, sum_over_ranks
    clips = list(range(7))
    mine = shard(clips, rank, world)
    ms = 10.0 + 5.0 * rank                       # rank 1 is slower
    worst = max_over_ranks(ms)
    total = sum_over_ranks(len(mine))
    dist.barrier()
    q.put((rank, mine, worst, total))
    dist.destroy_process_group()


def test_shard_and_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, w0, t0), (r1, m1, w1, t1) = res
    assert sorted(m0 + m1) == list(range(7)) and not set(m0) & set(m1)     # every clip exactly once
    assert w0 == w1 == 15.0                                                # everyone reports the slowest rank
    assert t0 == t1 == 7.0
