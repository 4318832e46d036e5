"""Rank layout of the measured runtime (runtime/cases.py::parallel_groups) under gloo: tensor-parallel ranks are adjacent, data-parallel
replicas stride by tp -- the order the analytical model assumes (tp fastest, then dp)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from simumax_b200.runtime.cases import Case, parallel_groups
        res = {}
        for tp in (1, 2, 4):
            case = Case(f"t{tp}", "llama3-8b", 2, 128, 1, 1, tp=tp, dp=world // tp)
            tp_group, dp_group = parallel_groups(case)
            members = {}
            for name, g in (("tp", tp_group), ("dp", dp_group)):
                if g is None:
                    members[name] = None
                    continue
                n = dist.get_world_size(g)
                got = [None] * n
                dist.all_gather_object(got, rank, group=g)
                members[name] = got
            res[tp] = members
        gathered = [None] * world
        dist.all_gather_object(gathered, res)
        if rank == 0:
            out.put(gathered)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(240)
def test_tp_ranks_adjacent_dp_ranks_strided():
    world = 4
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    per_rank = q.get(timeout=200)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, res in enumerate(per_rank):
        assert res[1] == {"tp": None, "dp": [0, 1, 2, 3]}                     # pure data parallel: the world is the dp group
        assert res[2]["tp"] == [rank // 2 * 2, rank // 2 * 2 + 1]             # tp pairs {0,1} {2,3}
        assert res[2]["dp"] == [rank % 2, rank % 2 + 2]                       # replicas {0,2} {1,3}
        assert res[4] == {"tp": None, "dp": None}                             # tp = world: the default group is the tp group
