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


# ------------------------------------------------------------------ stacked batch: one set of launches per IPM phase
def stack_models(models):
    """B independent QuadraticModels as ONE block-diagonal model (variables, rows, A and H of unit u shifted by the
    offsets off_n[u], off_m[u]); returns (stacked model, off_n, off_m)."""
    import numpy as np
    from .problems import QuadraticModel
    off_n = np.concatenate([[0], np.cumsum([q.nvar for q in models])]).astype(np.int64)
    off_m = np.concatenate([[0], np.cumsum([q.ncon for q in models])]).astype(np.int64)
    cat = lambda f: np.concatenate([f(q) for q in models]) if models else np.zeros(0)
    st = QuadraticModel(
        c=cat(lambda q: q.c),
        Hrows=np.concatenate([q.Hrows.astype(np.int64) + off_n[u] for u, q in enumerate(models)]),
        Hcols=np.concatenate([q.Hcols.astype(np.int64) + off_n[u] for u, q in enumerate(models)]),
        Hvals=cat(lambda q: q.Hvals),
        Arows=np.concatenate([q.Arows.astype(np.int64) + off_m[u] for u, q in enumerate(models)]),
        Acols=np.concatenate([q.Acols.astype(np.int64) + off_n[u] for u, q in enumerate(models)]),
        Avals=cat(lambda q: q.Avals),
        lcon=cat(lambda q: q.lcon), ucon=cat(lambda q: q.ucon), lvar=cat(lambda q: q.lvar), uvar=cat(lambda q: q.uvar),
        c0=0.0, x0=cat(lambda q: q.x0), y0=cat(lambda q: q.y0), name="stack_of_%d" % len(models))
    return st, off_n, off_m


class BatchedMPCSolver:
    """BASELINE config C5 as the reference would need it on a GPU: B independent LPs / QPs advance in lock-step through ONE
    set of kernel launches per IPM phase (the reference's mpc! loop, src/solver.jl:332-360, runs one problem at a time).
    The units are stacked into a block-diagonal problem for everything that is element-wise or linear algebra (assembly,
    the task-graph factorization of the elimination FOREST, solves, SpMVs); step lengths, centering parameter, barrier
    value, scaling and termination stay per unit (mipm_batch_*: one CTA per unit). Every unit's iterates are those of
    solving it alone (tests compare each unit with its own oracle trace). Default options only (AdaptiveStep /
    ConservativeStep, no Gondzio corrections), all-equality rows (no slack columns), no residual check."""

    def __init__(self, models, **kwargs):
        import numpy as np
        from . import solver as S
        self.models = list(models)
        self.B = len(self.models)
        if self.B < 1:
            raise ValueError("empty batch")
        if any(np.any(q.lcon != q.ucon) for q in self.models):
            raise NotImplementedError("batched path: all-equality rows only (no slack columns)")
        stacked, self.off_n, self.off_m = stack_models(self.models)
        kwargs = dict(kwargs)
        kwargs["fused"] = True
        self.s = S.MPCSolver(stacked, **kwargs)
        if not isinstance(self.s.opt.step_rule, (S.AdaptiveStep, S.ConservativeStep)) or self.s.opt.max_ncorr > 0:
            raise NotImplementedError("batched path: AdaptiveStep / ConservativeStep without Gondzio corrections")
        if self.s.opt.kkt_system not in ("Normal", "K2"):
            raise NotImplementedError("batched path: NormalKKTSystem or K2")
        self.s.h.batch_configure(self.off_n, self.off_m)
        self._seg = lambda a, off, u: a[off[u]:off[u + 1]]

    # ---- MadNLP.initialize! + set_scaling! with per-unit objective scaling
    def _initialize(self):
        import numpy as np
        import torch
        from . import solver as S
        s, h, opt = self.s, self.s.h, self.s.opt
        B, off_n = self.B, self.off_n
        H = s._host
        up_ = lambda t, a: t.copy_(a, non_blocking=True)
        up_(s.x, H["x0"]), up_(s.xl, H["xl"]), up_(s.xu, H["xu"]), up_(s.rhs, H["rhs"]), up_(s.cvec, H["c"])
        up_(s.y, H["y0"]), up_(s.A_V, H["A_V"])
        qp = s.qp
        if qp.nnzh > 0:
            up_(s.Hx, H["H_full"])
            if opt.kkt_system == "K2":
                up_(s.hess, H["H_tril"])
        h.init_bounds(s.n, opt.bound_relax_factor, opt.bound_push, opt.bound_fac, s.x, s.xl, s.xu)
        s.con_scale = np.ones(s.m)
        self.obj_scale = np.ones(B)
        if opt.scaling:
            if s._amax_A > 100.0:
                raise NotImplementedError("batched path: |A_ij| <= 100 (identity constraint scaling)")
            if qp.nnzh > 0:
                h.copy(s.n, s.cvec, s.f)
                h.hess_spmv(1.0, s.Hx, s.x, 1.0, s.f)
                gn = h.batch_amax(0, s.f)
            else:
                gn = h.batch_amax(0, s.cvec)
            with np.errstate(divide="ignore"):
                self.obj_scale = np.where(gn > 0, np.minimum(1.0, 100.0 / gn), 1.0)
        if np.any(self.obj_scale != 1.0):
            scale_n = torch.from_numpy(np.repeat(self.obj_scale, np.diff(off_n))).to(s.device)
            s.cvec.mul_(scale_n)                      # rare path: per-unit objective scaling on the stacked vector
            if qp.nnzh > 0:
                raise NotImplementedError("batched path: QP units need ||grad f(x0)|| <= 100 (identity objective scaling)")
        s.zl.zero_(), s.zu.zero_()
        self.norm_b = h.batch_amax(1, s.rhs) if s.m else np.zeros(B)
        h.fill(s.n, 0.0, s.jacl)
        h.fill(s.n, 1.0, s.reg), h.fill(s.n, 1.0, s.pr_diag), h.fill(s.m, 0.0, s.du_diag)
        h.fill(s.nlb, 0.0, s.l_lower), h.fill(s.nub, 0.0, s.u_lower)
        h.fill(s.nlb, 1.0, s.l_diag), h.fill(s.nub, 1.0, s.u_diag)
        s.init_regularization()
        s.compress_hessian()
        s.compress_jacobian()
        # objective, gradient and constraints at the pushed x0, per unit, BEFORE init_starting_point! moves x: the
        # reference does not re-evaluate them afterwards (stale values in iteration 0)
        self.obj_val = self.obj_scale * np.array([q.c0 for q in self.models]) + h.batch_dot(s.cvec, s.x)
        if qp.nnzh > 0:
            s.hH.hess_spmv(1.0, s.Hx, s.x, 0.0, s.buffer_n)
            self.obj_val = self.obj_val + 0.5 * h.batch_dot(s.buffer_n, s.x)
        s._eval_grad()
        s._eval_cons()
        self.norm_c = h.batch_amax(0, s.f)          # ||grad f(x0)||_inf per unit (quirk A.9 x)
        self._init_starting_point()
        s.jtprod(s.jacl, s.y)

    def _init_starting_point(self):
        """src/solver.jl:6-125 with the scalar stages per unit (mipm_batch_init_point_stage)."""
        import numpy as np
        s, h = self.s, self.s.h
        n, m = s.n, s.m
        N = s.p.numel()
        h.fill(n, s.del_w, s.reg), h.fill(n, s.del_w, s.pr_diag), h.fill(m, s.del_c, s.du_diag)
        s.factorize_wrapper()
        if not s.linear_solver.is_factorized():
            from .solver import SolveException
            raise SolveException("initial factorization failed")
        h.fill(N, 0.0, s.p)
        h.axpby(m, -1.0, s.c, 0.0, s.p[n:n + m])
        h.copy(N, s.p, s.d)
        s.kkt_solve(s.d)
        h.axpby(n, 1.0, s.d[:n], 1.0, s.x)
        h.fill(N, 0.0, s.p)
        h.axpby(n, -1.0, s.f, 0.0, s.p[:n])
        h.copy(N, s.p, s.d)
        s.kkt_solve(s.d)
        h.copy(m, s.d[n:n + m], s.y)
        s.jtprod(s.jacl, s.y)
        h.axpby(n, 1.0, s.f, 1.0, s.jacl)
        mins = h.batch_init_point_stage(0)
        delta_x = np.maximum(0.0, np.maximum(-1.5 * mins[:, 0], -1.5 * mins[:, 1]))
        delta_s = np.maximum(0.0, np.maximum(-1.5 * mins[:, 2], -1.5 * mins[:, 3]))
        st = h.batch_init_point_stage(1, delta_x, delta_s)
        with np.errstate(divide="ignore", invalid="ignore"):
            delta_x2 = st[:, 0] / (2 * (st[:, 1] + st[:, 2]))
            delta_s2 = st[:, 0] / (2 * (st[:, 3] + st[:, 4]))
        chk = h.batch_init_point_stage(2, delta_x2, delta_s2, s.opt.bound_fac)
        if s.nlb > 0 and not (np.all(chk[:, 0] > 0.0) and np.all(chk[:, 2] > 0.0)):
            raise AssertionError("starting point not strictly interior (lower)")

    def solve(self):
        """solve! for every unit; returns a list of ExecutionStats (one per unit, in order). The per-unit bookkeeping
        between the one synchronisation of an iteration and the launches of its second half is array arithmetic over
        the units (the GPU idles while it runs); traces are kept as per-iteration arrays and turned into the per-unit
        list of records on access; the result arrays are views of page-locked buffers owned by the solver (valid
        until the solve after next)."""
        import time
        import numpy as np
        import torch
        from . import solver as S
        s, h, opt, B = self.s, self.s.h, self.s.opt, self.B
        t0 = time.perf_counter()
        s.start_time = time.time()
        self._initialize()
        CODES = [S.REGULAR, S.SOLVE_SUCCEEDED, S.INFEASIBLE_PROBLEM_DETECTED, S.DIVERGING_ITERATES,
                 S.MAXIMUM_ITERATIONS_EXCEEDED, S.INTERNAL_ERROR]
        code = np.zeros(B, dtype=np.int64)               # index into CODES
        iters = np.zeros(B, dtype=np.int64)
        best = np.full(B, np.inf)
        active = np.ones(B, dtype=np.int32)
        alpha_p, alpha_d = np.zeros(B), np.zeros(B)
        mu = np.full(B, opt.mu_init)
        dobj_last = np.full(B, np.nan)
        c0 = self.obj_scale * np.array([q.c0 for q in self.models])
        nb1, nc1 = np.maximum(1.0, self.norm_b), np.maximum(1.0, self.norm_c)
        records = []                                     # per iteration: (k, del_w, active mask, columns)
        started = False
        k = 0
        worst = np.full(B, np.inf)
        s.del_w_trace = s.del_w
        while True:
            trace_del_w = s.del_w
            s.update_regularization()
            act = active == 1
            # every unit still running is within 10 x tol: read the measures before the next factorization goes out, so
            # the pass in which the last units converge does not pay for a factorization of the whole stack
            peek = started and bool(np.all(worst[act] <= 10.0 * opt.tol))
            ok = True
            if peek:
                out = h.batch_peek()
            else:
                out, ok = h.batch_iter_begin(s.del_w, s.del_c)
            if started:
                self.obj_val = np.where(act, c0 + out[:, 5] + 0.5 * out[:, 6], self.obj_val)
                alpha_p = np.where(act, out[:, 7], alpha_p)
                alpha_d = np.where(act, out[:, 8], alpha_d)
                mu = np.where(act, out[:, 9], mu)
            dobj, dnorm = out[:, 0], out[:, 4]
            inf_pr, inf_du, inf_compl = out[:, 1] / nb1, out[:, 2] / nc1, out[:, 3] / nc1
            with np.errstate(invalid="ignore"):
                best = np.where(act & (inf_compl < best), inf_compl, best)          # min(best, inf_compl)
                dobj_last = np.where(act, dobj, dobj_last)
                worst = np.where(inf_du > inf_pr, inf_du, inf_pr)                    # max(inf_pr, inf_du, inf_compl) with the
                worst = np.where(inf_compl > worst, inf_compl, worst)                # comparison order of the scalar code
                a_obj = np.abs(self.obj_val)
                thr_inf = np.where(1.0 > 10.0 * a_obj, 1.0, 10.0 * a_obj)             # max(10 |obj|, 1)
                a_dobj = np.abs(dobj)
                thr_div = np.where(a_dobj > 10.0, a_dobj, 10.0)
                new_code = np.select(
                    [worst <= opt.tol,
                     (inf_compl > opt.divergence_tol * best) & (dobj > thr_inf),
                     self.obj_val < -opt.divergence_tol * thr_div,
                     np.full(B, k >= opt.max_iter),
                     ~np.isfinite(worst)],
                    [1, 2, 3, 4, 5], default=0)
            records.append((k, trace_del_w, act.copy(), self.obj_val / self.obj_scale, dobj / self.obj_scale, inf_pr, inf_du,
                            inf_compl, mu.copy(), alpha_p.copy(), alpha_d.copy(), np.zeros(B) if k == 0 else dnorm.copy()))
            done_now = act & (new_code != 0)
            code = np.where(done_now, new_code, code)
            iters = np.where(done_now, k, iters)
            active = np.where(done_now, 0, active).astype(np.int32)
            if not active.any():
                break
            if peek:
                _, ok = h.batch_iter_begin(s.del_w, s.del_c)
            h.batch_set_active(active)
            for _ in range(2):                       # factorize_regularized_system! retries, for the whole stack
                if ok:
                    break
                s.del_w *= 100.0
                s.del_c *= 100.0
                ok = h.mpc_refactor(s.del_w, s.del_c)
            rule = opt.step_rule
            ir = max(opt.ir_steps, 1 if s._has_free else 0)       # free variables: refine every solve (like MPCSolver)
            if isinstance(rule, S.AdaptiveStep):
                h.batch_iter_rest(opt.mu_min, 0, rule.tau_min, ir)
            else:
                h.batch_iter_rest(opt.mu_min, 1, rule.tau, ir)
            started = True
            k += 1
        torch.cuda.synchronize(s.device)
        total = time.perf_counter() - t0
        h.spmv(0, 1.0, s.AT_x, s.x, 0.0, s.buffer_m)
        n, m = s.n, s.m
        s._res_turn ^= 1
        slab = s._res_slabs[s._res_turn]
        parts, off = [], 0
        for src, ln in ((s.x, n), (s.buffer_m, m), (s.y, m), (s.zl, n), (s.zu, n)):
            dst = slab[off:off + ln]
            if ln:
                dst.copy_(src, non_blocking=True)
            parts.append(dst.numpy())
            off += ln
        torch.cuda.synchronize(s.device)
        x, cons, y, zl, zu = parts
        if np.any(self.obj_scale != 1.0):
            y /= np.repeat(self.obj_scale, np.diff(self.off_m))
            sn = np.repeat(self.obj_scale, np.diff(self.off_n))
            zl /= sn
            zu /= sn
        counters = dict(launches=h.launch_count(), iterations_of_the_batch=k, ls_stats=s.linear_solver.stats)
        obj = self.obj_val / self.obj_scale
        dob = dobj_last / self.obj_scale
        res = []
        for u in range(B):
            n0, n1, m0, m1 = self.off_n[u], self.off_n[u + 1], self.off_m[u], self.off_m[u + 1]
            res.append(S.ExecutionStats(
                status=CODES[code[u]], iter=int(iters[u]), objective=obj[u], dual_objective=dob[u],
                solution=x[n0:n1], constraints=cons[m0:m1], multipliers=y[m0:m1], multipliers_L=zl[n0:n1],
                multipliers_U=zu[n0:n1], trace=_UnitTrace(records, u), total_time=total, linear_solver_time=0.0,
                counters=counters))
        return res


class _UnitTrace:
    """Trace of one unit of a batch (a sequence of the same records MPCSolver keeps), materialised on first access from
    the per-iteration arrays of the whole batch."""
    _KEYS = ("objective", "dual_objective", "inf_pr", "inf_du", "inf_compl", "mu", "alpha_p", "alpha_d", "dnorm")

    def __init__(self, records, u):
        self._records, self._u, self._list = records, u, None

    def _get(self):
        if self._list is None:
            u, out = self._u, []
            for rec in self._records:
                k, del_w, act = rec[0], rec[1], rec[2]
                if not act[u]:
                    break
                d = dict(k=k, del_w=del_w)
                for name, col in zip(self._KEYS, rec[3:]):
                    d[name] = float(col[u])
                out.append(d)
            self._list = out
        return self._list

    def __len__(self):
        return len(self._get())

    def __getitem__(self, i):
        return self._get()[i]

    def __iter__(self):
        return iter(self._get())


def solve_batch_stacked(models, **kwargs):
    """madipm for a batch of independent models through the stacked path; returns one ExecutionStats per model."""
    return BatchedMPCSolver(models, **kwargs).solve()
