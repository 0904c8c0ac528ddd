"""The N>1 path on CPU (gloo, world_size 2): static instance sharding, max-over-ranks timing, result gathering,
and that every rank derives the same cached analysis for the same pattern (replicas only, no data-path collective)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sleqp_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sleqp_b200 import Symbolic, problems

    mine = shard.assign_instances(9, world, rank)
    hashes = {}
    for i in mine:
        p = problems.poisson_control(10, 2, seed=100 + i)  # same pattern size, different active sets / data
        hashes[i] = Symbolic(p.N, *p.kkt_lower()).stats()["perm_hash"]
    same = problems.poisson_control(10, 2, seed=7)
    common = Symbolic(same.N, *same.kkt_lower()).stats()["perm_hash"]
    slowest = shard.max_over_ranks(1.0 + rank, dist)
    everything = shard.gather_objects((rank, mine, hashes, common), dist)
    dist.barrier()
    if rank == 0:
        q.put((slowest, everything))
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    slowest, everything = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert slowest == 2.0  # max over ranks
    covered = sorted(i for _, mine, _, _ in everything for i in mine)
    assert covered == list(range(9))  # every instance exactly once
    assert everything[0][1] == [0, 2, 4, 6, 8] and everything[1][1] == [1, 3, 5, 7]
    assert everything[0][3] == everything[1][3]  # same pattern -> same analysis on every rank


def test_assignment_edge_cases():
    assert shard.assign_instances(3, 8, 5) == []
    assert shard.assign_instances(64, 8, 7) == list(range(7, 64, 8))
    assert sum(len(shard.assign_instances(64, w, r)) for w in (4,) for r in range(4)) == 64
    assert shard.max_over_ranks(3.5) == 3.5 and shard.gather_objects("x") == ["x"]
