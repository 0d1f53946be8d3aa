"""Sharding of a batch of independent atoms over G ranks (one process per GPU, no collective: SURVEY §8e).

Longest-processing-time-first assignment with cost = number of (spin) orbitals of the atom; heavy atoms
are placed first so that the slowest chains start early on every rank."""
from typing import List, Sequence

from .api import aufbau, split_spin


def atom_cost(Z: int, method: int = 0) -> int:
    if method:
        a, b, _, _ = split_spin(Z)
        return len(a) + len(b)
    return len(aufbau(Z))


def partition_atoms(Zs: Sequence[int], n_ranks: int, method: int = 0) -> List[List[int]]:
    """Returns, per rank, the list of indices into Zs it owns (deterministic)."""
    order = sorted(range(len(Zs)), key=lambda i: (-atom_cost(Zs[i], method), i))
    loads = [0] * n_ranks
    parts: List[List[int]] = [[] for _ in range(n_ranks)]
    for i in order:
        r = min(range(n_ranks), key=lambda k: (loads[k], k))
        parts[r].append(i)
        loads[r] += atom_cost(Zs[i], method)
    return [sorted(p) for p in parts]
