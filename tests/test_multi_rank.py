"""World-size-2 gloo tests (CPU) of the multi-rank host logic: unit sharding for the batch
config (C5) and the max-over-ranks timing reduction bench.py uses. The GPU solve itself is
replaced by the CPU oracle here -- this tests the plumbing, not the kernels."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from madipm_jl_b200.batch import shard_range, solve_batch
from madipm_jl_b200.problems import random_sparse_lp


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 128, 1024, 1025):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.mpc_oracle import madipm as oracle_madipm

    def make(i):
        return random_sparse_lp(20, 60, 3, 100 + i, structure="uniform")

    res = solve_batch(make, 5, solve_fn=lambda qp: oracle_madipm(qp, kkt_system="Normal"))
    # the bench's timing reduction: max over ranks
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out.put((rank, [(r.index, r.status, r.iter, round(r.objective, 9), r.rank) for r in res], float(t.item())))
    dist.destroy_process_group()


def test_solve_batch_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    (r0, res0, t0), (r1, res1, t1) = got
    assert res0 == res1, "every rank must see the same gathered result"
    assert [x[0] for x in res0] == list(range(5))
    assert all(x[1] == "SOLVE_SUCCEEDED" for x in res0)
    assert [x[4] for x in res0] == [0, 0, 0, 1, 1]        # 3 units on rank 0, 2 on rank 1
    assert t0 == t1 == 2.0
    # sharded result == unsharded result
    from oracle.mpc_oracle import madipm as oracle_madipm
    single = [round(oracle_madipm(random_sparse_lp(20, 60, 3, 100 + i, structure="uniform"), kkt_system="Normal").objective, 9)
              for i in range(5)]
    assert [x[3] for x in res0] == single
