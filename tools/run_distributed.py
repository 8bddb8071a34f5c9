"""Distributed (block-angular, config C4) runs under torchrun: parity against the oracle on a small instance, then
timing of the C4-scale instance. Usage:
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_distributed.py [scale]"""
import json, os, sys, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from madipm_jl_b200.problems import block_angular_lp, config_c4
from madipm_jl_b200.solver import MPCSolver

def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # ---- parity on a small instance (every rank runs the oracle redundantly: it is tiny)
    from oracle.mpc_oracle import madipm as oracle_madipm
    qp = block_angular_lp(6, 9, 8, 12, 5)
    ref = oracle_madipm(qp, kkt_system="Normal")
    got = MPCSolver(qp, kkt_system="Normal", linear_solver="distributed", n_border=qp.meta["n_border"], device=local).solve()
    dev = max(abs(a[f] - b[f]) / max(1.0, abs(a[f]), abs(b[f])) for a, b in zip(got.trace, ref.trace)
              for f in ("objective", "dual_objective", "inf_pr", "inf_du", "inf_compl"))
    ok = got.status == ref.status and abs(got.iter - ref.iter) <= 2 and dev <= 1e-8
    if rank == 0:
        print(json.dumps({"check": "parity_small", "ranks": world, "status": got.status, "iters": got.iter, "oracle_iters": ref.iter,
                          "max_trace_dev": dev, "ok": bool(ok)}), flush=True)
    # ---- C4-scale timing
    qp = config_c4(scale=scale)
    t0 = time.time()
    s = MPCSolver(qp, kkt_system="Normal", linear_solver="distributed", n_border=qp.meta["n_border"], device=local)
    t_setup = time.time() - t0
    s.solve()                                   # warm-up solve
    s.k = 0; s.trace = []
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    r = s.solve()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t1
    tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        st = s.linear_solver.stats
        print(json.dumps({"config": "C4 scale %g" % scale, "ranks": world, "m": qp.ncon, "n": qp.nvar, "n_border": qp.meta["n_border"],
                          "components": s.linear_solver.n_components, "status": r.status, "iterations": r.iter, "objective": r.objective,
                          "solve_s": float(tt.item()), "iters_per_s": r.iter / float(tt.item()), "setup_s": t_setup,
                          "local_symbolic": {k: st[k] for k in ("n", "nnz_l", "flops", "n_supernodes", "n_levels", "max_front_cols")}}), flush=True)
    if os.environ.get("MIPM_DIST_PROFILE"):
        # coarse per-stage wall times (synchronised; perturbs the run, so it is done after the timed solve)
        ls = s.linear_solver
        acc = {}
        def timed(name, fn):
            def w(*a, **k):
                torch.cuda.synchronize(); t = time.perf_counter()
                out = fn(*a, **k)
                torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
                acc[name + "_n"] = acc.get(name + "_n", 0) + 1
                return out
            return w
        h = ls.h
        orig_stage, orig_sstage, orig_ar = h.ls_factorize_stage, h.ls_solve_stage, ls._allreduce
        h.ls_factorize_stage = lambda nz, st: timed("factor_stage%d" % st, orig_stage)(nz, st)
        h.ls_solve_stage = lambda b, st: timed("solve_stage%d" % st, orig_sstage)(b, st)
        ls._allreduce = lambda t, op=None: timed("allreduce_%d" % t.numel(), orig_ar)(t, op)
        ls.factorize = timed("factorize_total", ls.factorize)
        ls.solve = timed("ls_solve_total", ls.solve)
        s.k = 0; s.trace = []
        torch.cuda.synchronize(); t1 = time.perf_counter()
        r = s.solve()
        torch.cuda.synchronize(); acc["solve_s"] = time.perf_counter() - t1
        if rank == 0:
            print(json.dumps({"profile": {k: (round(v, 5) if isinstance(v, float) else v) for k, v in sorted(acc.items())},
                              "iterations": r.iter, "ranks": world}), flush=True)
    if world > 1:
        dist.destroy_process_group()

if __name__ == "__main__":
    main()
