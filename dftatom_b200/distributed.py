"""Multi-GPU driver: one process per GPU, atoms sharded over ranks, results gathered on every rank.

Atoms are independent: there is no collective on the data path.  The only communication is the final
gather of the per-atom result records (a few hundred bytes per atom), done with torch.distributed's object gather
(NCCL on the GPU box, gloo in the CPU tests)."""
from typing import Callable, List, Sequence

from .shard import partition_atoms


def solve_sharded(options: Sequence, solve_fn: Callable[[list], list], rank: int, world: int, gather_fn=None) -> List:
    """Solve `options` (list of Options) cooperatively.

    solve_fn(list_of_options) -> list_of_results runs this rank's shard (normally Context.solve_batch).
    gather_fn(obj) -> list over ranks of obj (default: torch.distributed.all_gather_object when world > 1).
    Returns the results of ALL atoms, in the order of `options`, on every rank."""
    parts = partition_atoms([o.Z for o in options], world, method=max(o.method for o in options))
    mine = parts[rank]
    local = solve_fn([options[i] for i in mine]) if mine else []
    payload = list(zip(mine, local))
    if world == 1:
        gathered = [payload]
    elif gather_fn is not None:
        gathered = gather_fn(payload)
    else:
        import torch.distributed as dist
        gathered = [None] * world
        dist.all_gather_object(gathered, payload)
    out = [None] * len(options)
    for part in gathered:
        for i, r in part:
            out[i] = r
    assert all(r is not None for r in out), "a shard is missing"
    return out
