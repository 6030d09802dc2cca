"""world_size-2 gloo tests (CPU) of the N>1 host logic: sharding, max-over-ranks timing, result gather."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qiskit_addon_sqd_b200._dispatch import gather_results, max_over_ranks, shard_indices


def test_shards_are_disjoint_and_cover():
    for n in (0, 1, 5, 8, 17):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                seen += shard_indices(n, r, world)
            assert sorted(seen) == list(range(n))
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def _worker(rank, world, port, n_units, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard_indices(n_units, rank, world)
        # every rank "solves" its own units; the result carries the unit id so the order is checkable
        local = [{"unit": k, "energy": -1.0 * k, "rank": rank} for k in mine]
        t = max_over_ranks(10.0 + rank)            # slowest rank defines the step time
        everything = gather_results(local, n_units)
        dist.barrier()
        out_q.put((rank, mine, t, [r["unit"] for r in everything], [r["rank"] for r in everything]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, n_units = 2, 5
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    assert got[0][1] == [0, 2, 4] and got[1][1] == [1, 3]
    for rank, mine, t, units, owners in got:
        assert t == 11.0                              # max over ranks, identical everywhere
        assert units == list(range(n_units))          # gathered back in unit order
        assert owners == [k % world for k in range(n_units)]
    assert torch.cuda.is_available() or True
