"""Population sharding: one process per GPU, individuals distributed round-robin, fitness gathered on the host.

The reference shards a generation the same way over MPI ranks (optimization/program.py:534-538:
``if i % number_of_mpi_processes == mpi_rank``) and exchanges pickled fitness values with
``allgather`` (:285-291).  There is no data-path collective: the only communication is the gather of
``(time, convergence factor, iterations)`` tuples, done with ``torch.distributed`` object collectives
(NCCL or gloo process group; tiny messages)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

Fitness = Tuple[float, float, float]


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Indices evaluated by ``rank``: i % world == rank (program.py:535)."""
    return [i for i in range(n_items) if i % world == rank]


def merge_shards(shards: Sequence[Sequence], n_items: int) -> list:
    """Inverse of :func:`shard_indices`: shards[r][k] is item r + k*world."""
    world = len(shards)
    out: list = [None] * n_items
    for r, shard in enumerate(shards):
        for k, v in enumerate(shard):
            out[r + k * world] = v
    if any(v is None for v in out):
        raise ValueError("incomplete shards")
    return out


def evaluate_sharded(items: Sequence, evaluate_local: Callable[[Sequence], List[Fitness]], rank: int = 0,
                     world: int = 1, dist=None) -> List[Fitness]:
    """Every rank evaluates its shard with ``evaluate_local`` and receives the complete, ordered result list
    (the reference's allgather semantics)."""
    mine = [items[i] for i in shard_indices(len(items), rank, world)]
    local = list(evaluate_local(mine))
    if len(local) != len(mine):
        raise ValueError("evaluate_local returned a wrong number of results")
    if world == 1 or dist is None:
        return local
    gathered: List[Optional[list]] = [None] * world
    dist.all_gather_object(gathered, local)
    return merge_shards(gathered, len(items))
