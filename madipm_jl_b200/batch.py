"""Sharding of independent LP/QP instances across ranks (BASELINE config C5).

Independent units shard with no data-path collective: rank r solves the contiguous slice
`shard_range(n_units, r, world)` on its own GPU with its own library handles; the only
communication is the final gather of per-unit statistics (torch.distributed, NCCL on GPUs,
gloo in the CPU tests of the host logic)."""
from dataclasses import asdict, dataclass


def shard_range(n_units, rank, world):
    """Contiguous, balanced partition of range(n_units): sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


@dataclass
class UnitResult:
    index: int
    status: str
    iter: int
    objective: float
    rank: int


def _gpu_solve(qp, **kwargs):
    from .solver import madipm
    return madipm(qp, **kwargs)


def solve_batch(make_model, n_units, solve_fn=None, **kwargs):
    """Solve units make_model(0..n_units-1), sharded over the ranks of the default process group
    (or alone when torch.distributed is not initialised). Returns the full, index-ordered list
    of UnitResult on every rank."""
    import torch.distributed as dist
    solve_fn = solve_fn or _gpu_solve
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    lo, hi = shard_range(n_units, rank, world)
    mine = []
    for i in range(lo, hi):
        st = solve_fn(make_model(i), **kwargs)
        mine.append(asdict(UnitResult(i, st.status, int(st.iter), float(st.objective), rank)))
    if world == 1:
        return [UnitResult(**d) for d in mine]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    flat = sorted((d for part in gathered for d in part), key=lambda d: d["index"])
    return [UnitResult(**d) for d in flat]
