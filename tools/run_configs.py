"""Runs the BASELINE.json configs other than the bench workload through the CUDA path and prints one
JSON line per config (status, iterations, time, symbolic stats). Usage:
  python tools/run_configs.py c1 c3 c5 [--c3-scale S] [--c5-units N] [--c5-threads T --c5-grid G]
Under torchrun, C5 units are sharded across ranks (madipm_jl_b200.batch.solve_batch)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madipm_jl_b200 import problems  # noqa: E402
from madipm_jl_b200.batch import solve_batch  # noqa: E402
from madipm_jl_b200.solver import MPCSolver  # noqa: E402


def run_one(name, qp, kkt, **kw):
    t0 = time.time()
    s = MPCSolver(qp, kkt_system=kkt, **kw)
    t_setup = time.time() - t0
    s.solve()                       # untimed warm-up solve (module load, cooperative-launch setup), like bench.py
    s.k = 0
    s.trace = []
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    r = s.solve()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t1
    st = s.linear_solver.stats
    print(json.dumps({"config": name, "kkt": kkt, "status": r.status, "iterations": r.iter, "objective": r.objective,
                      "dual_objective": r.dual_objective, "solve_s": dt, "iters_per_s": r.iter / dt, "setup_s": t_setup,
                      "refinements": r.counters.get("refinements", 0),
                      "symbolic": {k: st[k] for k in ("n", "nnz_a", "nnz_l", "flops", "n_supernodes", "n_levels", "max_front_cols")}}),
          flush=True)
    return r


def main():
    args = sys.argv[1:]
    opt = {"--c3-scale": 1.0, "--c5-units": 32, "--c5-threads": 12, "--c5-grid": 32}
    for k in list(opt):
        if k in args:
            i = args.index(k)
            opt[k] = float(args[i + 1])
            del args[i:i + 2]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if "c1" in args and rank == 0:
        qp = problems.config_c1()
        run_one("C1", qp, "Normal", device=local)
        run_one("C1", qp, "K2", device=local)
    if "c3" in args and rank == 0:
        qp = problems.config_c3(scale=opt["--c3-scale"])
        run_one("C3 scale %g" % opt["--c3-scale"], qp, "K2", device=local)
    if "c5" in args:
        n_units = int(opt["--c5-units"])
        from madipm_jl_b200.batch import shard_range
        lo, hi = shard_range(n_units, rank, world)
        models = {i: problems.config_c5(i) for i in range(lo, hi)}      # the synthetic generator is not part of the timing
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = solve_batch(lambda i: models[i], n_units, kkt_system="Normal", device=local,
                          threads=int(opt["--c5-threads"]), grid_limit=int(opt["--c5-grid"]))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if rank == 0:
            ok = sum(r.status == "SOLVE_SUCCEEDED" for r in res)
            print(json.dumps({"config": "C5", "units": n_units, "ranks": world, "threads_per_gpu": int(opt["--c5-threads"]),
                              "grid_limit": int(opt["--c5-grid"]), "succeeded": ok, "wall_s": dt,
                              "lps_per_s": n_units / dt, "mean_iters": float(np.mean([r.iter for r in res])),
                              "units_per_rank": [sum(r.rank == q for r in res) for q in range(world)]}), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
