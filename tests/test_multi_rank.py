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


def test_block_angular_partition_reproduces_the_global_schur_complement():
    """Host logic of the distributed solver (numpy only): interior blocks are split over ranks, every rank's
    local system keeps the border last, and the sum of the ranks' Schur contributions on the border equals the
    Schur complement of the global matrix -- which is what the NCCL all-reduce of the root panel assembles."""
    import numpy as np
    import scipy.sparse as sp
    from madipm_jl_b200 import _lib
    from madipm_jl_b200.distributed import local_system, partition_interior
    from madipm_jl_b200.problems import block_angular_lp
    from oracle import sparse_ref
    qp = block_angular_lp(5, 6, 5, 9, 3)
    m, n, nb = qp.ncon, qp.nvar, qp.meta["n_border"]
    Bp, Bj, Bm = _lib.coo_to_csr(m, n, qp.Arows, qp.Acols)
    Cp, Cj = sparse_ref.build_normal_system(m, n, Bp, Bj)
    D = np.random.default_rng(0).uniform(0.5, 2.0, n)
    Cx = sparse_ref.assemble_normal_system(m, n, Bp, Bj, qp.Avals[Bm], Cp, Cj, D)
    low = sp.csc_matrix((Cx, Cj, Cp), shape=(m, m))
    C = (low + sp.tril(low, -1).T).toarray()
    ni = m - nb
    S_ref = C[ni:, ni:] - C[ni:, :ni] @ np.linalg.solve(C[:ni, :ni], C[:ni, ni:])
    for world in (1, 2, 3):
        owner, ncomp = partition_interior(m, Cp, Cj, nb, world)
        assert ncomp == 5 and len(owner) == ni
        counts = np.bincount(owner, minlength=world)
        assert counts.max() - counts.min() <= (ni // 5) + 1               # whole blocks, balanced
        S_sum = np.zeros((nb, nb))
        seen = np.zeros(ni, dtype=int)
        nz_ext = np.concatenate([Cx, [0.0]])
        for r in range(world):
            loc, cp, ri, nzmap = local_system(m, Cp, Cj, nb, owner, r)
            assert list(loc[-nb:]) == list(range(ni, m)) and np.all(np.diff(loc) > 0)
            seen[loc[:-nb]] += 1
            Ll = sp.csc_matrix((nz_ext[nzmap], ri, cp), shape=(len(loc), len(loc)))
            Cl = (Ll + sp.tril(Ll, -1).T).toarray()
            k = len(loc) - nb
            S_sum += Cl[k:, k:] - (Cl[k:, :k] @ np.linalg.solve(Cl[:k, :k], Cl[:k, k:]) if k else 0.0)
            # the library's border analysis keeps the border as the final supernode of the local system
            h = _lib.Handle(device=-1)
            h.ls_analyze_border(len(loc), cp, ri, nb)
            sym = h.ls_symbolic()
            assert sym["sn_ptr"][-2] == k and list(sym["perm"][-nb:]) == list(range(k, len(loc)))
        assert np.all(seen == 1)
        assert np.abs(S_sum - S_ref).max() <= 1e-10 * np.abs(S_ref).max()
