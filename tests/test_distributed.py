"""CPU test of the N > 1 host path: world_size-2 gloo processes shard a batch, solve their shards and gather.
The per-rank solver here is the CPU oracle (tests may use it); on the GPU box the same driver wraps Context.solve_batch."""
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import dftatom_b200 as D
    import oracle_lib as O
    from dftatom_b200.distributed import solve_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    opts = [D.Options(Z, 9, 12.0, 0.008, 0.5, 0) for Z in (1, 2, 3, 4, 10, 6)]

    def solve(shard):
        return [O.scf(o.Z, o.MultigridLevels, o.alpha, o.MaxR, o.deltaGrid, o.method, max_vcycles=10)["steps"][-1]["Etotal"] for o in shard]

    res = solve_sharded(opts, solve, rank, world)
    if rank == 0:
        ret["sharded"] = res
        ret["serial"] = solve(opts)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["sharded"] == ret["serial"]          # bit-identical: no cross-atom arithmetic exists
    assert len(ret["sharded"]) == 6
