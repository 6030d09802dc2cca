"""Multi-GPU plumbing for the K independent subspaces of an SQD iteration.

The reference runs the K diagonalisations of ``solve_sci_batch`` one after the other
(``fermion.py:670-681``) and notes that they are embarrassingly parallel (``README.md:76``).  Here the
units are sharded, never the data: subspace k goes to rank ``k % world`` (one process per GPU under
``torchrun``) or to device ``k % n_devices`` (single process, ``solve_sci_batch(devices=...)``).
No collective touches the data path; ``torch.distributed`` is used only for the barrier and for the
max-over-ranks reduction of timings, and to gather the (KB-sized) results when a caller wants them on
every rank.
"""

from __future__ import annotations

from typing import Any, Sequence


def shard_indices(n_units: int, rank: int, world: int) -> list[int]:
    """Indices of the units rank ``rank`` owns: round-robin, disjoint, covering."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    return list(range(rank, n_units, world))


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar (timings); identity when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_results(local: Sequence[Any], n_units: int) -> list[Any]:
    """All-gather per-rank result lists (python objects) back into unit order on every rank."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return list(local)
    world, rank = dist.get_world_size(), dist.get_rank()
    bucket: list[Any] = [None] * world
    dist.all_gather_object(bucket, list(local))
    out: list[Any] = [None] * n_units
    for r in range(world):
        for j, k in enumerate(shard_indices(n_units, r, world)):
            out[k] = bucket[r][j]
    del rank
    return out
