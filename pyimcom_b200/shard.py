"""Multi-GPU sharding of the coaddition path: one process per GPU, blocks (or strips of 2x2 stamp groups of one
block) are independent, so there is no data-path collective; the only exchange is the final gather of the output
cube on rank 0 (the reference itself only launches one OS process per block, docs/run_README.rst:78-100).

Works with any torch.distributed backend: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def assign_blocks(n_blocks: int, world: int, rank: int) -> list[int]:
    """Static round-robin of block indices over ranks (stamps per block are constant, so static is balanced)."""
    return list(range(rank, n_blocks, world))


def assign_stamp_groups(n1P: int, world: int, rank: int):
    """Single-block sharding: contiguous strips of whole 2x2 stamp-group rows (keeps group-anchored PSFs local,
    coadd.py:2056-2060).  Returns the OutStamp (j, i) list of this rank in the reference's traversal order."""
    rows = list(range(1, n1P + 1, 2))
    per = (len(rows) + world - 1) // world
    mine = rows[rank * per:(rank + 1) * per]
    out = []
    for j in mine:
        for i in range(1, n1P + 1, 2):
            for dj in range(2):
                for di in range(2):
                    if j + dj <= n1P and i + di <= n1P:
                        out.append((j + dj, i + di))
    return out


def gather_cube(local: torch.Tensor, world: int, rank: int, dst: int = 0):
    """Gather every rank's output cube on ``dst``: returns the stacked (world, ...) tensor there, None elsewhere."""
    if world == 1:
        return local.unsqueeze(0)
    local = local.contiguous()
    if rank == dst:
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.gather(local, gather_list=parts, dst=dst)
        return torch.stack(parts)
    dist.gather(local, gather_list=None, dst=dst)
    return None


def reduce_cube(local: torch.Tensor, world: int, dst: int = 0):
    """Sum-reduce zero-initialised full cubes (single block split into strips: the fade-border overlap-add across
    strip seams is done by the reduction)."""
    if world > 1:
        dist.reduce(local, dst=dst, op=dist.ReduceOp.SUM)
    return local
