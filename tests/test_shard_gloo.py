"""Multi-rank host logic on CPU: world_size 2 over gloo (the GPU box runs the same code over NCCL).

Blocks and postage stamps are independent, so ranks never exchange data on the hot path; the only collective is the
final gather (or sum-reduce, when one block is split into strips) of the output cube."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyimcom_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n1P, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. every rank coadds its own blocks; rank 0 receives the stacked cubes in rank order
        blocks = shard.assign_blocks(5, world, rank)
        local = torch.full((1, 2, 6, 6), float(rank + 1)) * (1 + len(blocks))
        got = shard.gather_cube(local, world, rank)
        # 2. one block split into strips of 2x2 stamp groups: zero-initialised full cubes, sum-reduced
        mine = shard.assign_stamp_groups(n1P, world, rank)
        side = n1P * 4 + 2
        cube = torch.zeros((1, 1, side, side))
        for (j, i) in mine:  # each stamp adds 1 to its (4+2)^2 footprint: seams overlap-add across ranks
            cube[0, 0, (j - 1) * 4:(j - 1) * 4 + 6, (i - 1) * 4:(i - 1) * 4 + 6] += 1.0
        red = shard.reduce_cube(cube, world)
        if rank == 0:
            q.put(("gather", got.numpy().copy(), blocks))
            q.put(("reduce", red.numpy().copy(), mine))
        else:
            assert got is None
            q.put(("mine", None, (blocks, mine)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gather_and_strip_reduce():
    world, n1P = 2, 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n1P, q)) for r in range(world)]
    for p in procs:
        p.start()
    items = [q.get(timeout=120) for _ in range(3)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = {k: (a, b) for k, a, b in items}
    g, blocks0 = res["gather"]
    blocks1, mine1 = res["mine"][1]
    assert sorted(blocks0 + blocks1) == list(range(5)) and not set(blocks0) & set(blocks1)
    assert g.shape == (2, 1, 2, 6, 6)
    assert np.all(g[0] == 1.0 * (1 + len(blocks0))) and np.all(g[1] == 2.0 * (1 + len(blocks1)))
    red, mine0 = res["reduce"]
    assert sorted(mine0 + mine1) == [(j, i) for j in range(1, n1P + 1) for i in range(1, n1P + 1)]
    ref = np.zeros_like(red)
    for j in range(1, n1P + 1):
        for i in range(1, n1P + 1):
            ref[0, 0, (j - 1) * 4:(j - 1) * 4 + 6, (i - 1) * 4:(i - 1) * 4 + 6] += 1.0
    assert np.array_equal(red, ref)


def test_assignments_are_partitions():
    for world in (1, 2, 3, 4, 8):
        allb = sum((shard.assign_blocks(11, world, r) for r in range(world)), [])
        assert sorted(allb) == list(range(11))
        for n1P in (4, 6, 84):
            alls = sum((shard.assign_stamp_groups(n1P, world, r) for r in range(world)), [])
            assert len(alls) == n1P * n1P == len(set(alls))
