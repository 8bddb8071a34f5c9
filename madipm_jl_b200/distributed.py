"""Distributed direct solver for block-angular normal matrices (BASELINE config C4).

New functionality (the reference has no multi-GPU code): the elimination tree of a block-angular
A D A' is a forest of independent interior blocks (one per commodity) under one dense root separator
(the linking constraints). Each rank factors the subtrees of ITS blocks; the only exchange per
factorization is one NCCL all-reduce of the root Schur block, and per solve one all-reduce of the
root right-hand side plus one of the assembled solution. Everything else in the iteration (vector
kernels, SpMV, assembly) runs replicated on every rank, so MPCSolver is unchanged above the
linear-solver interface (factorize / is_factorized / solve, like MadNLP.AbstractLinearSolver).

The host logic here (partition of the interior blocks over ranks, index maps) is pure numpy and is
tested on CPU with gloo; the numeric work goes through the staged C-ABI entries
mipm_ls_analyze_border / mipm_ls_factorize_stage / mipm_ls_solve_stage.
"""
import numpy as np
import scipy.sparse as sp
from scipy.sparse.csgraph import connected_components


def partition_interior(n, colptr, rowval, n_border, world):
    """Connected components of the interior graph (vertices < n - n_border), assigned to ranks by
    decreasing size onto the least-loaded rank (deterministic). Returns owner[n_interior] -> rank."""
    ni = n - n_border
    low = sp.csc_matrix((np.ones(len(rowval), dtype=np.int8), rowval, colptr), shape=(n, n))
    inner = low[:ni, :ni]
    ncomp, label = connected_components(inner + inner.T, directed=False)
    sizes = np.bincount(label, minlength=ncomp)
    order = np.lexsort((np.arange(ncomp), -sizes))           # by size descending, ties by label
    load = np.zeros(world, dtype=np.int64)
    comp_owner = np.zeros(ncomp, dtype=np.int64)
    for c in order:
        r = int(np.argmin(load))                              # first least-loaded rank
        comp_owner[c] = r
        load[r] += sizes[c]
    return comp_owner[label], ncomp


def local_system(n, colptr, rowval, n_border, owner, rank):
    """Local lower-CSC pattern of rank `rank`: its interior vertices (ascending) then the border.
    Returns (loc2glob, colptr_loc, rowval_loc, nzmap) where nzmap indexes the GLOBAL nzval; border x border
    entries on ranks != 0 point at the extra zero slot len(rowval) (the caller keeps a trailing 0.0 there)."""
    ni = n - n_border
    mine = np.flatnonzero(owner == rank)
    loc = np.concatenate([mine, np.arange(ni, n)]).astype(np.int64)
    low = sp.csc_matrix((np.arange(1, len(rowval) + 1, dtype=np.float64), rowval, colptr), shape=(n, n))
    sub = low[loc][:, loc].tocsc()
    sub.sort_indices()
    nzmap = (sub.data - 1).astype(np.int64)
    if rank != 0:
        nloc_i = len(mine)
        cols = np.repeat(np.arange(len(loc)), np.diff(sub.indptr))
        bb = (sub.indices >= nloc_i) & (cols >= nloc_i)
        nzmap[bb] = len(rowval)
    return loc, sub.indptr.astype(np.int32), sub.indices.astype(np.int32), nzmap


class _DevView:
    """Zero-copy torch view of library-owned device memory (through __cuda_array_interface__)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class DistributedB200Solver:
    def __init__(self, n, colptr, rowval, nzval, n_border, device, stream, group=None):
        import torch
        import torch.distributed as dist
        from . import _lib
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        assert nzval.numel() == len(rowval) + 1, "global nzval needs one trailing zero slot"
        self.n, self.n_border, self.nzval = n, n_border, nzval
        owner, self.n_components = partition_interior(n, colptr, rowval, n_border, self.world)
        loc, cp, ri, nzmap = local_system(n, colptr, rowval, n_border, owner, self.rank)
        self.n_loc, self.n_int = len(loc), len(loc) - n_border
        dev = torch.device("cuda", device)
        self.d_loc = torch.from_numpy(loc).to(dev)
        self.d_nzmap = torch.from_numpy(nzmap).to(dev)
        self.nz_loc = torch.zeros(max(len(ri), 1), dtype=torch.float64, device=dev)
        self.b_loc = torch.zeros(self.n_loc, dtype=torch.float64, device=dev)
        self.flag = torch.zeros(1, dtype=torch.float64, device=dev)
        self.h = _lib.Handle(device=device, stream=stream)
        self.h.ls_analyze_border(self.n_loc, cp, ri, n_border)
        self.stats = self.h.ls_stats()
        proot, nroot, prhs = self.h.ls_root_info()
        assert nroot == n_border
        self.root = torch.as_tensor(_DevView(proot, (n_border * n_border,)), device=dev)
        self.root_rhs = torch.as_tensor(_DevView(prhs, (n_border,)), device=dev)
        self.nnz_loc = len(ri)

    def _allreduce(self, t, op=None):
        if self.world > 1:
            self.dist.all_reduce(t, op=op or self.dist.ReduceOp.SUM, group=self.group)

    def factorize(self):
        self.h.gather(self.nnz_loc, self.nzval, self.d_nzmap, self.nz_loc)
        self.h.ls_factorize_stage(self.nz_loc, 0)        # subtrees below the root; root panel = local Schur part
        self._allreduce(self.root)                       # THE exchange step: sum of the Schur contributions
        self.h.ls_factorize_stage(self.nz_loc, 1)        # global root front, redundantly on every rank

    def is_factorized(self):
        ok = self.h.ls_status()
        self.flag.fill_(0.0 if ok else 1.0)
        self._allreduce(self.flag, self.dist.ReduceOp.MAX if self.world > 1 else None)
        return float(self.flag.item()) == 0.0

    def solve(self, x):
        h = self.h
        h.gather(self.n_loc, x, self.d_loc, self.b_loc)
        if self.rank != 0:
            h.fill(self.n_border, 0.0, self.b_loc[self.n_int:])      # the border RHS enters once, on rank 0
        h.ls_solve_stage(self.b_loc, 0)
        self._allreduce(self.root_rhs)
        h.ls_solve_stage(self.b_loc, 1)
        if self.world == 1:
            h.scatter(self.n_loc, self.b_loc, self.d_loc, x)        # one rank owns every row: no exchange
            return x
        h.fill(self.n, 0.0, x)
        cnt = self.n_loc if self.rank == 0 else self.n_int          # border solution is identical everywhere
        h.scatter(cnt, self.b_loc, self.d_loc, x)
        self._allreduce(x)
        return x

    def inertia(self):
        return self.h.ls_inertia()

    def introduce(self):
        return "madipm_b200 distributed supernodal Cholesky (%d ranks, root separator %d)" % (self.world, self.n_border)
