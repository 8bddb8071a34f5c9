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


def _gpu_solve_concurrent(make_model, indices, threads, grid_limit, **kwargs):
    """Units of one rank solved `threads` at a time on ONE GPU: every worker thread has its own CUDA stream and its own
    library handles, and the persistent kernels of a handle are capped at `grid_limit` CTAs (mipm_set_grid_limit) so the
    cooperative launches of different units are resident together. A small LP cannot fill a B200 (a 500 x 500 normal
    matrix is 8 block steps of a handful of tiles) and an IPM iteration is a chain of dependent launches, so units are
    overlapped instead. The C library releases the GIL in every call (ctypes).
    Measured on one B200 (128 C5 units, tools/run_configs.py): 2.0 s one at a time, 0.50-0.68 s with 12 threads and
    grid_limit 32 (five runs). Earlier multi-second stalls came from per-handle cudaMallocHost / cudaFreeHost /
    cudaGetDeviceProperties calls, which synchronise the device; the library now caches those per process."""
    import queue
    import threading
    import torch
    from .solver import MPCSolver
    todo = queue.Queue()
    for i in indices:
        todo.put(i)
    out, errors, lock = {}, [], threading.Lock()
    device = kwargs.get("device", 0)

    def worker():
        torch.cuda.set_device(device)
        stream = torch.cuda.Stream(device=device)
        with torch.cuda.stream(stream):
            while True:
                try:
                    i = todo.get_nowait()
                except queue.Empty:
                    return
                try:
                    st = MPCSolver(make_model(i), grid_limit=grid_limit, **kwargs).solve()
                    with lock:
                        out[i] = st
                except Exception as exc:       # surfaced by the caller: a failed unit must not look solved
                    with lock:
                        errors.append((i, exc))
                    return

    ts = [threading.Thread(target=worker) for _ in range(max(1, min(threads, len(indices))))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errors:
        raise RuntimeError("unit %d failed: %r" % errors[0])
    return [(i, out[i]) for i in indices]


def solve_batch(make_model, n_units, solve_fn=None, threads=1, grid_limit=0, **kwargs):
    """Solve units make_model(0..n_units-1), sharded over the ranks of the default process group
    (or alone when torch.distributed is not initialised). Returns the full, index-ordered list
    of UnitResult on every rank. threads > 1 (GPU path only) overlaps that many units per GPU, each on its own
    stream with its persistent kernels capped at grid_limit CTAs (use e.g. threads=12, grid_limit=32 for small LPs)."""
    import torch.distributed as dist
    solve_fn = solve_fn or _gpu_solve
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    lo, hi = shard_range(n_units, rank, world)
    mine = []
    if solve_fn is _gpu_solve and threads > 1 and hi > lo:
        for i, st in _gpu_solve_concurrent(make_model, list(range(lo, hi)), threads, grid_limit, **kwargs):
            mine.append(asdict(UnitResult(i, st.status, int(st.iter), float(st.objective), rank)))
    else:
        for i in range(lo, hi):
            st = solve_fn(make_model(i), **kwargs)
            mine.append(asdict(UnitResult(i, st.status, int(st.iter), float(st.objective), rank)))
    if world == 1:
        return [UnitResult(**d) for d in mine]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    flat = sorted((d for part in gathered for d in part), key=lambda d: d["index"])
    return [UnitResult(**d) for d in flat]
