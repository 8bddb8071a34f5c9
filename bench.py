#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric   : IPM iterations per second (plus time-to-1e-8 and factorization FP64 TFLOP/s as extra keys)
workload : config C2 = synthetic sparse LP, m=200,000 constraints, n=1,000,000 variables, 8 nnz/col,
           rows drawn within +-50 of each column's home row (uniformly random rows make chol(AA')
           ~90% dense = 147 GB at this size: SURVEY fact 8), NormalKKTSystem + supernodal Cholesky.
step     : one pass of the mpc! loop body (src/solver.jl:333-359): termination test, KKT assembly,
           numeric factorization, predictor + corrector solves (each with the residual check of
           solve_system!; no refinement round is needed on this workload), ratio tests, step,
           model re-evaluation.
extras   : after the headline measurement the same run also reports (key "extras"): BASELINE configs[4] (1 024
           independent LPs through the stacked batch path, sharded over the ranks), configs[3] (the block-angular
           LP with the distributed solver over all ranks: the one config with an exchange step), the mesh variant
           of C2 with fronts of thousands of columns (rank 0, N = 1) and a second CPU point (SciPy SuperLU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--scale S]

N > 1: the single-instance configs do not shard ("replicas only"), so every rank runs an
independent replica of the workload; value = total iterations of all ranks / max-over-ranks time.
`--impl reference` times the CPU restatement of the reference (oracle/, numpy + scalar C LDL^T on the
host cores) on the same config: the Julia reference itself cannot run in this image (no Julia).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# dram__bytes_read.sum + dram__bytes_write.sum of one k_factor_persistent launch from `ncu --set full`
# (profiles/ncu_factor_persistent_r01_final_raw.csv), keyed by --scale; None when not captured for that size
TRAFFIC_BYTES = {}
try:
    _tp = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", f) for f in ("factor_traffic_r02.json", "factor_traffic_r01.json")]
    for _k, _v in json.load(open([f for f in _tp if os.path.exists(f)][0])).items():
        try:
            TRAFFIC_BYTES[float(_k)] = float(_v)
        except (TypeError, ValueError):
            pass            # descriptive entries (source, capture details)
except Exception:
    pass

SOLVE_TRAFFIC = {}
try:
    for _k, _v in json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "solve_traffic_r02.json"))).items():
        try:
            SOLVE_TRAFFIC[float(_k)] = float(_v)
        except (TypeError, ValueError):
            pass
except Exception:
    pass

METRIC = "ipm_iterations_per_second"
UNIT = "iter/s"


def load_peaks():
    peaks = {"hbm_gbs": 6650.0, "hbm_src": "fallback", "fp64_tflops": 35.46, "fp64_src": "own DGEMM measurement (profiles/fp64_peak_r01.json)"}
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks["hbm_gbs"], peaks["hbm_src"] = float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    try:
        p = json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r01.json")))
        peaks["fp64_tflops"] = float(p["fp64_tflops"])
    except Exception:
        pass
    return peaks


def make_workload(scale):
    from madipm_jl_b200.problems import config_c2
    qp = config_c2(seed=2, scale=scale)
    name = "C2: LP m=%d n=%d 8 nnz/col, rows within +-50 of home row, NormalKKT + Cholesky" % (qp.ncon, qp.nvar)
    return qp, name


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world, out):
    """CPU restatement of the reference on the host cores (kind = "port")."""
    if rank != 0:
        return
    from oracle.mpc_oracle import MPCOracle
    qp, wname = make_workload(args.scale)
    t0 = time.time()
    o = MPCOracle(qp, kkt_system="Normal", linear_solver="ldl", fast_symbolic=True)
    o.start_time = time.time()
    o.initialize()
    t_setup = time.time() - t0
    budget = 200.0
    for _ in range(args.warmup):
        if not o.mpc_iteration():
            o.initialize()
    done = 0
    t1 = time.perf_counter()
    while done < args.steps and (time.perf_counter() - t1) < budget:
        if not o.mpc_iteration():
            tpause = time.perf_counter()
            o.initialize()
            t1 += time.perf_counter() - tpause
            continue
        done += 1
    dt = time.perf_counter() - t1
    val = done / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(done, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wname, "kkt_system": "Normal", "linear_solver": "oracle LDL^T (Davis up-looking, RCM), 1 thread"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": "%d full-size IPM iterations of the CPU restatement (oracle/), setup %.1fs untimed" % (done, t_setup)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "Julia reference cannot run in this image (no Julia); cuDSS not installed",
    }
    print(json.dumps(line), file=out, flush=True)


def run_extras(args, rank, world, local_rank, peaks, barrier):
    """The other BASELINE configs in the same run (never allowed to break the headline line)."""
    import torch
    import torch.distributed as dist
    out = {}

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- C5: 1 024 independent LPs (m=500, n=2 000), sharded over the ranks, each shard one stacked batch
    try:
        from madipm_jl_b200.batch import BatchedMPCSolver, shard_range
        from madipm_jl_b200.problems import config_c5
        n_units = max(world, int(1024 * min(1.0, args.scale)))
        lo, hi = shard_range(n_units, rank, world)
        t0 = time.perf_counter()
        b = BatchedMPCSolver([config_c5(i) for i in range(lo, hi)], kkt_system="Normal", device=local_rank)
        torch.cuda.synchronize()
        t_ctor = time.perf_counter() - t0
        b.solve()                                   # warm-up
        barrier()
        t1 = time.perf_counter()
        res = b.solve()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t1)
        ok = sum(r.status == "SOLVE_SUCCEEDED" for r in res)
        okt = torch.tensor([ok], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(okt)
        out["c5"] = {"workload": "C5: %d independent LPs m=500 n=2000 5 nnz/col, %d per rank, stacked batch path" % (n_units, hi - lo),
                     "units": n_units, "solve_s": dt, "lps_per_s": n_units / dt, "succeeded": int(okt.item()),
                     "construct_s_rank0": t_ctor, "batch_iterations_rank0": res[0].counters["iterations_of_the_batch"],
                     "scaling": "weak-free (independent units, no collective)"}
        del b, res
        torch.cuda.empty_cache()
    except Exception as exc:      # noqa: BLE001
        out["c5"] = {"error": repr(exc)}

    # ---- C4: block-angular multicommodity-flow LP, distributed solver over all ranks (strong scaling)
    try:
        from madipm_jl_b200.problems import config_c4
        from madipm_jl_b200.solver import MPCSolver
        qp = config_c4(scale=args.scale)
        t0 = time.perf_counter()
        s4 = MPCSolver(qp, kkt_system="Normal", linear_solver="distributed", n_border=qp.meta["n_border"], device=local_rank)
        torch.cuda.synchronize()
        t_ctor = time.perf_counter() - t0
        s4.solve()                                  # warm-up
        s4.k = 0
        s4.trace = []
        barrier()
        t1 = time.perf_counter()
        r4 = s4.solve()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t1)
        st4 = s4.linear_solver.stats
        out["c4"] = {"workload": "C4: block-angular multicommodity flow m=%d n=%d, %d linking rows, distributed solver on %d rank(s)"
                                 % (qp.ncon, qp.nvar, qp.meta["n_border"], world),
                     "status": r4.status, "iterations": r4.iter, "solve_s": dt, "iter_per_s": r4.iter / dt, "objective": r4.objective,
                     "construct_s_rank0": t_ctor, "scaling": "strong (one LP over all ranks, all-reduce of the root Schur block)",
                     "local_symbolic_rank0": {k_: st4[k_] for k_ in ("n", "nnz_l", "flops", "n_supernodes", "n_levels", "max_front_cols")}}
        del s4, r4, qp
        torch.cuda.empty_cache()
    except Exception as exc:      # noqa: BLE001
        out["c4"] = {"error": repr(exc)}

    if world == 1 and rank == 0:
        # ---- mesh variant of C2: fronts with thousands of columns (tensor-pipe bound updates)
        try:
            from madipm_jl_b200.problems import config_c2_mesh
            from madipm_jl_b200.solver import MPCSolver
            qp = config_c2_mesh(scale=args.scale)
            sm = MPCSolver(qp, kkt_system="Normal", device=local_rank)
            sm.solve()
            sm.k = 0
            sm.trace = []
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            rm = sm.solve()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t1
            pf = sm.h.ls_factorize_profile(sm.aug_nz)
            stm = sm.linear_solver.stats
            kms = pf["kernel"]["ms"]
            out["c2_mesh"] = {"workload": "mesh variant of C2: LP m=%d n=%d 8 nnz/col, rows within +-3 cells on a 2-D mesh" % (qp.ncon, qp.nvar),
                              "status": rm.status, "iterations": rm.iter, "solve_s": dt, "iter_per_s": rm.iter / dt,
                              "symbolic": {k_: stm[k_] for k_ in ("n", "nnz_l", "flops", "n_supernodes", "n_levels", "max_front_cols", "max_front_rows")},
                              "roofline": {"kernel": "k_factor_tasks", "bound": "tensor", "achieved": stm["flops"] / kms / 1e9,
                                           "peak": peaks["fp64_tflops"], "unit": "TFLOP/s", "frac": stm["flops"] / kms / 1e9 / peaks["fp64_tflops"],
                                           "avg_launch_ms": kms,
                                           "update_tasks_tflops": pf["update"]["work"] / pf["update"]["ms"] / 1e9,
                                           "update_tasks_frac": pf["update"]["work"] / pf["update"]["ms"] / 1e9 / peaks["fp64_tflops"]},
                              "factor_classes": pf}
            del sm, rm, qp
            torch.cuda.empty_cache()
        except Exception as exc:      # noqa: BLE001
            out["c2_mesh"] = {"error": repr(exc)}
        # ---- second CPU point: the same restatement with SciPy SuperLU as the linear solver
        if not args.no_cpu_baseline:
            try:
                from oracle.mpc_oracle import MPCOracle
                qp, _ = make_workload(args.scale)
                tb = time.time()
                o = MPCOracle(qp, kkt_system="Normal", linear_solver="splu", fast_symbolic=True)
                o.start_time = time.time()
                o.initialize()
                nit, tc = 0, time.perf_counter()
                while nit < 2 and o.mpc_iteration():
                    nit += 1
                dtc = time.perf_counter() - tc
                out["cpu_superlu"] = {"value": nit / dtc, "unit": UNIT, "cores": 1, "kind": "port",
                                      "sample": "%d full-size C2 iterations of the CPU restatement with scipy.sparse.linalg.splu (SuperLU, COLAMD) as "
                                                "linear solver; setup %.1fs untimed; no multithreaded sparse direct solver exists in this image"
                                                % (nit, time.time() - tb - dtc), "timers_s": o.timers}
            except Exception as exc:      # noqa: BLE001
                out["cpu_superlu"] = {"error": repr(exc)}
    return out


def _claim_stdout():
    """Keep stdout for the ONE JSON line: native libraries (NCCL prints its version banner with
    printf) get fd 1 redirected to stderr; the returned file object writes to the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="scale m and n of the workload (testing only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C4 / C5 / mesh / SuperLU blocks")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from madipm_jl_b200.solver import MPCSolver
    peaks = load_peaks()
    qp, wname = make_workload(args.scale)
    torch.zeros(1, device="cuda").item()          # CUDA context creation (0.1-1.2 s on a fresh box) is not part of the setup time
    t0 = time.time()
    solver = MPCSolver(qp, kkt_system="Normal", device=local_rank)
    t_setup = time.time() - t0
    setup_log = [[n_, round(t_, 4)] for n_, t_ in solver.setup_log]
    st = solver.linear_solver.stats
    h = solver.h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- e2e: the public solve call with host buffers (H2D of the model, D2H of the result)
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_first = time.perf_counter()
    solver.solve()          # warm-up solve: first-use costs (module load, cooperative launch setup); reported as "cold"
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t_first
    solver.k = 0
    solver.trace = []
    barrier()
    t1 = time.perf_counter()
    res = solver.solve()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t1
    iters_total = res.iter
    n, m = solver.n, solver.m
    h2d = solver.h2d_bytes_per_solve          # pinned problem data uploaded by every solve (solver._stage_problem)
    d2h = 8 * (3 * n + 2 * m)               # x, zl, zu, y and the constraint values A x: what solve() copies into its result buffers
    if world > 1:
        tt = torch.tensor([t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e_max = float(tt.item())
    else:
        t_e2e_max = t_e2e
    e2e_val = world * iters_total / t_e2e_max

    # ---------------- device-timed steps
    solver.k = 0
    solver.trace = []
    solver.start_time = time.time()
    solver.initialize()
    for _ in range(args.warmup):
        if not solver.mpc_iteration():
            solver.k = 0
            solver.initialize()
    barrier()
    l0 = h.launch_count()
    done, total_ms = 0, 0.0
    while done < args.steps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        alive = True
        while done < args.steps and alive:
            alive = solver.mpc_iteration()
            if alive:
                done += 1
        e1.record()
        torch.cuda.synchronize()
        total_ms += e0.elapsed_time(e1)
        if not alive:   # converged inside the timed region: restart from the starting point (untimed)
            solver.k = 0
            solver.initialize()
    barrier()
    launches = h.launch_count() - l0
    # keep the same workload running (untimed) until nvidia-smi has delivered a few samples
    t_extra = time.time()
    while len(sampler.rows) < 5 and time.time() - t_extra < 3.0:
        if not solver.mpc_iteration():
            solver.k = 0
            solver.initialize()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    if world > 1:
        tt = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    value = world * args.steps / (total_ms * 1e-3)

    # ---------------- per-stage device timing and rooflines (rank 0 reports)
    def time_call(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    counts = np.bincount(solver.Aj, minlength=n).astype(np.float64)
    T = float(np.sum(counts * (counts + 1) / 2))
    nnzC = len(solver.aug_rowval)
    asm_bytes = 12.0 * T + 12.0 * nnzC + 8.0 * n
    asm_ms = time_call(lambda: h.normal_assemble(solver.pr_diag, solver.aug_nz, False))
    prof = h.ls_factorize_profile(solver.aug_nz)
    fac_ms = time_call(lambda: h.ls_factorize_async(solver.aug_nz), reps=3)
    xb = torch.randn(m, dtype=torch.float64, device="cuda")
    sol_ms = time_call(lambda: h.ls_solve(xb, 0), reps=3)
    sol_bytes = 2.0 * (8.0 * st["nnz_l"] + 4.0 * (st["nnz_l"] ** 0.5))  # L read twice (fwd + bwd)
    spmv_ms = time_call(lambda: h.spmv(0, 1.0, solver.AT_x, solver.x, 0.0, solver.buffer_m))
    spmv_bytes = 12.0 * len(solver.Aj) + 8.0 * (n + m)
    stages = {
        "assembly": {"ms": asm_ms, "bound": "hbm", "achieved_gbs": asm_bytes / asm_ms / 1e6,
                     "frac": asm_bytes / asm_ms / 1e6 / peaks["hbm_gbs"], "algorithmic_bytes": asm_bytes},
        "factorization": {"ms": fac_ms, "tflops": st["flops"] / fac_ms / 1e9, "flops": st["flops"],
                          "launches": st["n_launches"]},
        "triangular_solve_pair": {"ms": sol_ms, "bound": "hbm", "achieved_gbs": sol_bytes / sol_ms / 1e6,
                                  "frac": sol_bytes / sol_ms / 1e6 / peaks["hbm_gbs"], "algorithmic_bytes": sol_bytes,
                                  # DRAM bytes of one launch from the ncu capture (the sweeps also stream the inverted
                                  # diagonal blocks, which the algorithmic figure does not count)
                                  "traffic": SOLVE_TRAFFIC.get(args.scale),
                                  "dram_gbs": (SOLVE_TRAFFIC[args.scale] / sol_ms / 1e6) if args.scale in SOLVE_TRAFFIC else None,
                                  "dram_frac": (SOLVE_TRAFFIC[args.scale] / sol_ms / 1e6 / peaks["hbm_gbs"]) if args.scale in SOLVE_TRAFFIC else None},
        "spmv": {"ms": spmv_ms, "bound": "hbm", "achieved_gbs": spmv_bytes / spmv_ms / 1e6,
                 "frac": spmv_bytes / spmv_ms / 1e6 / peaks["hbm_gbs"]},
        "factor_classes": prof,
    }
    # Dominant kernel of the step: k_factor_persistent (one cooperative launch = one numeric factorization,
    # ~60% of the step). Its FLOPs go through the FP64 tensor pipe (DMMA), so the bound is "tensor"; achieved =
    # algorithmic flops of the factorization (sum_j colcount_j^2) / average launch duration (CUDA events).
    kernel_ms = prof["kernel"]["ms"]           # span of the task kernel (first task start .. last task end, %globaltimer)
    achieved = st["flops"] / kernel_ms / 1e9
    upd = prof["update"]
    roof = {"kernel": "k_factor_tasks", "bound": "tensor", "achieved": achieved, "peak": peaks["fp64_tflops"],
            "unit": "TFLOP/s", "frac": achieved / peaks["fp64_tflops"], "traffic": TRAFFIC_BYTES.get(args.scale),
            "peak_source": "FP64 " + peaks["fp64_src"] + "; MEASURED_PEAKS.json has no FP64 entry",
            "avg_launch_ms": kernel_ms, "algorithmic_flops_per_launch": st["flops"],
            "share_of_step": kernel_ms / ms_per_step,
            "update_tasks": {"tflops": upd["work"] / upd["ms"] / 1e9, "frac": upd["work"] / upd["ms"] / 1e9 / peaks["fp64_tflops"],
                             "cta_ms": upd["ms"], "algorithmic_flops": upd["work"],
                             "what": "algorithmic update flops / (busy time of the trailing-update tasks summed over CTAs / grid): "
                                     "the rate the tensor-pipe tasks sustain while they run"},
            "note": "avg_launch_ms is the span of the task kernel from its own %globaltimer stamps (CUDA-event time of the whole "
                    "factorization incl. zero-fill + scatter is stages.factorization.ms); stages.factor_classes gives the CTA-busy "
                    "time per task class and the dependency wait; the kernel is bound by the critical path of the elimination tree "
                    "(profiles/r02_notes.md), not by the tensor pipe or HBM"}

    # ---------------- CPU baseline: oracle on a bounded sample of the same workload (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.mpc_oracle import MPCOracle
        tb = time.time()
        o = MPCOracle(qp, kkt_system="Normal", linear_solver="ldl", fast_symbolic=True)
        o.start_time = time.time()
        o.initialize()
        nit = 0
        tc = time.perf_counter()
        while nit < 3 and o.mpc_iteration():
            nit += 1
        dtc = time.perf_counter() - tc
        cpu = {"value": nit / dtc, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "%d full-size IPM iterations of the CPU restatement (numpy + scalar C LDL^T, RCM ordering, "
                         "nnz(L)=%d); setup %.1fs untimed; host has %d cores" % (nit, o.ls.nnzL, time.time() - tb - dtc, os.cpu_count()),
               "timers_s": o.timers}

    opt_ir, opt_tol = solver.opt.ir_steps, solver.opt.tol
    extras = None
    if not args.no_extras:
        del solver, h
        torch.cuda.empty_cache()
        extras = run_extras(args, rank, world, local_rank, peaks, barrier)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname, "kkt_system": "NormalKKTSystem", "linear_solver": "supernodal Cholesky (own)",
                       "ordering": "nested dissection", "ir_steps": opt_ir, "tol": opt_tol,
                       "l2": "inputs larger than L2 (L+U = %.0f MB)" % (8e-6 * (st["nnz_l"] + st["update_doubles"])),
                       "parallelism": "replicas x%d" % world, "scale": args.scale},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d / max(iters_total, 1),
                    "d2h_bytes_per_step": d2h / max(iters_total, 1), "iterations": iters_total,
                    "time_to_tol_s": t_e2e_max, "status": res.status, "initialize_s": res.counters.get("initialize_time"),
                    "refinements": res.counters.get("refinements", 0),
                    "what": "MPCSolver.solve(): host model -> H2D, init_starting_point!, mpc! to tol, D2H of x,y,zl,zu"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "time_to_1e-8_s": t_e2e_max, "iterations_to_1e-8": iters_total,
            "factor_fp64_tflops": st["flops"] / fac_ms / 1e9,
            "stages": stages,
            "symbolic": {k_: st[k_] for k_ in ("n", "nnz_a", "nnz_l", "nnz_l_exact", "flops", "n_supernodes", "n_levels",
                                              "max_front_cols", "max_front_rows", "update_doubles", "n_launches")},
            "setup_s": t_setup, "final_objective": res.objective,
            "cold": {"construct_s": t_setup, "first_solve_s": t_first, "cold_time_to_1e-8_s": t_setup + t_first,
                     "setup_log": setup_log,
                     "what": "MPCSolver(qp) (symbolic analysis on host and device, uploads; CUDA context already created) + the first solve() of the process"},
            "extras": extras,
        }
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
