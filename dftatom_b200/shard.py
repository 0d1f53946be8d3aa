"""Sharding of a batch of independent atoms over G ranks (one process per GPU, no collective: SURVEY §8e).

Longest-processing-time-first assignment on the library's cost model (dftatom_estimate_cost: orbitals x expected SCF steps, so that
the 60-150-step atoms - Cu, Zn, Ho..Yb - land on different ranks and weigh what they cost).  The same C function shards the batch of
`bin/dftatom --gpus N`."""
from typing import List, Sequence

from .api import estimate_cost, partition


def atom_cost(Z: int, method: int = 0) -> float:
    return estimate_cost(Z, method)


def partition_atoms(Zs: Sequence[int], n_ranks: int, method=0) -> List[List[int]]:
    """Returns, per rank, the list of indices into Zs it owns (deterministic).  method: one int or one per atom."""
    methods = [method] * len(Zs) if isinstance(method, int) else list(method)
    rank_of = partition(list(Zs), methods, n_ranks)
    return [[i for i, r in enumerate(rank_of) if r == k] for k in range(n_ranks)]
