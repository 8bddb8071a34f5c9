"""Host-side mirror of MadIPM's solver API on top of the C-ABI CUDA library.

Same names, argument meaning and error behaviour as the reference's Julia host code, which in
a Julia deployment stays Julia and reaches the same C ABI through `ccall`:

    MPCSolver(qp; kwargs...)        src/structure.jl:79-178
    solve!(solver) / madipm(qp)     src/solver.jl:362-428
    initialize!, init_starting_point!, mpc! and its steps   src/solver.jl:6-360
    factorize_regularized_system!, solve_system!            src/linear_solver.jl:6-44
    NormalKKTSystem / SparseKKTSystem (K2)                   src/KKT/normalkkt.jl, MadNLP

Everything numeric runs in libmadipm_b200.so on the GPU; this file only sequences calls and does
host scalar logic (status tests, regularization schedule, barrier update). torch is used for
device allocations and host<->device copies only. There is no CPU fallback.
"""
import time
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from ._lib import MpcModel, MpcVectors

# MadNLP.Status names
INITIAL = "INITIAL"
REGULAR = "REGULAR"
SOLVE_SUCCEEDED = "SOLVE_SUCCEEDED"
INFEASIBLE_PROBLEM_DETECTED = "INFEASIBLE_PROBLEM_DETECTED"
DIVERGING_ITERATES = "DIVERGING_ITERATES"
MAXIMUM_ITERATIONS_EXCEEDED = "MAXIMUM_ITERATIONS_EXCEEDED"
MAXIMUM_WALLTIME_EXCEEDED = "MAXIMUM_WALLTIME_EXCEEDED"
INTERNAL_ERROR = "INTERNAL_ERROR"
ERROR_IN_STEP_COMPUTATION = "ERROR_IN_STEP_COMPUTATION"


class SolveException(RuntimeError):
    """MadNLP.SolveException (thrown by solve_system!, src/linear_solver.jl:40-42)."""


# ---- option types (src/utils.jl:17-48)
@dataclass
class ConservativeStep:
    tau: float = 0.995


@dataclass
class AdaptiveStep:
    tau_min: float = 0.99


@dataclass
class MehrotraAdaptiveStep:
    gamma_f: float = 0.99


class NoRegularization:
    pass


@dataclass
class FixedRegularization:
    delta_p: float
    delta_d: float


@dataclass
class AdaptiveRegularization:
    delta_p: float
    delta_d: float
    delta_min: float


@dataclass
class IPMOptions:
    """src/utils.jl:69-105. kkt_system: "K2" (MadNLP.SparseKKTSystem, the default), "K2.5" (MadNLP.ScaledSparseKKTSystem)
    or "Normal" (MadIPM.NormalKKTSystem). linear-solver options of the B200 solver: ordering, ir_steps."""
    tol: float = 1e-8
    kkt_system: str = "K2"
    max_iter: int = 3000
    grid_limit: int = 0            # cap on the persistent kernels' grid (0 = whole GPU); see batch.solve_batch
    max_wall_time: float = 1e6
    divergence_tol: float = 1e4
    scaling: bool = True
    bound_push: float = 1e-2
    bound_fac: float = 1e-2
    bound_relax_factor: float = 1e-12
    regularization: object = field(default_factory=lambda: FixedRegularization(1e-10, 1e-10))
    step_rule: object = field(default_factory=lambda: AdaptiveStep(0.99))
    max_ncorr: int = 0
    mu_init: float = 1e-1
    mu_min: float = 1e-12
    tol_linear_solve: float = 1e-8
    check_residual: bool = False
    rethrow_error: bool = False
    # B200Solver options (the analogue of cudss_algorithm / ir options of MadNLPGPU.CUDSSSolver)
    ordering: int = _lib.MIPM_ORDER_ND
    # cudss_algorithm of MadNLPGPU.CUDSSSolver (test/test_gpu.jl:9-16 uses CHOLESKY for the normal equations,
    # scripts/benchmarks_gpu.jl:42 LDL): "auto" = "CHOLESKY" for NormalKKTSystem, "LDL" for K2 / K2.5. "LDL" with the
    # normal equations runs the square-root-free kernels: a pivot that rounding made non-positive late in an
    # ill-conditioned solve is then not a breakdown (no x100 regularization retry), as with the oracle's LDL'.
    cudss_algorithm: str = "auto"
    ir_steps: int = 0             # refinement rounds inside every linear solve (on the reduced system)
    # adaptive refinement on the full unreduced KKT system, driven by the residual that
    # solve_system! computes anyway (src/linear_solver.jl:29-35): refine while
    # ||p - K d||_inf / max(1, ||p||_inf) > refine_tol, at most max_refine times
    refine_tol: float = 1e-9
    max_refine: int = 3
    # one host synchronisation per iteration: step lengths, centering parameter and barrier value stay on the
    # device (mipm_mpc_iter_begin / mipm_mpc_iter_rest). Used for the default options (no Gondzio corrections,
    # Adaptive / Conservative step rule); otherwise the fine-grained host-driven sequence runs.
    fused: bool = True
    # "b200": the single-GPU supernodal solver; "distributed": block-angular normal matrix whose last n_border
    # rows are the linking constraints, factored over the ranks of the default process group (distributed.py)
    linear_solver: str = "b200"
    n_border: int = 0
    exact_assembly_order: bool = False
    device: int = 0
    # 0: C / Python conventions. 1: every index array crosses the C ABI exactly as the Julia glue passes it (1-based
    # Int32 patterns, 1-based Int64 value maps and bound indices, `index_base = 1` in every call): ext/MadIPMB200Ext
    index_base: int = 0


@dataclass
class ExecutionStats:
    """MadNLP.MadNLPExecutionStats subset."""
    status: str
    iter: int
    objective: float
    dual_objective: float
    solution: np.ndarray
    constraints: np.ndarray
    multipliers: np.ndarray
    multipliers_L: np.ndarray
    multipliers_U: np.ndarray
    trace: list
    total_time: float
    linear_solver_time: float
    counters: dict


def _dev(a, device, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device)


class B200Solver:
    """The MadNLP.AbstractLinearSolver mirror (constructor / factorize! / solve! / is_factorized /
    inertia / introduce), replacing MadNLPGPU.CUDSSSolver. Works on the lower-triangular CSC
    `aug_com` handed over by the KKT system."""

    def __init__(self, handle, n, colptr, rowval, nzval, kind, ordering, ir_steps, index_base=0):
        self.h, self.nzval, self.ir_steps = handle, nzval, ir_steps
        self.kind = kind
        handle.ls_analyze(n, colptr, rowval, kind=kind, ordering=ordering, index_base=index_base)
        self.stats = handle.ls_stats()

    def factorize(self):
        self.h.ls_factorize_async(self.nzval)

    def is_factorized(self):
        return self.h.ls_status()

    def solve(self, x):
        self.h.ls_solve(x, self.ir_steps)
        return x

    def inertia(self):
        return self.h.ls_inertia()

    def introduce(self):
        return "madipm_b200 supernodal %s" % ("Cholesky" if self.kind == _lib.MIPM_CHOLESKY else "LDL^T")


class MPCSolver:
    def __init__(self, qp, **kwargs):
        t0 = time.time()
        self.setup_log = []          # (stage, seconds since the start of the constructor): where construction time goes
        _mark = lambda name: self.setup_log.append((name, time.time() - t0))
        self.opt = IPMOptions(**kwargs)
        opt = self.opt
        if not torch.cuda.is_available():
            raise RuntimeError("madipm_jl_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", opt.device)
        torch.cuda.set_device(self.device)
        self.qp = qp
        # ---- MadNLP.get_index_constraints (App. B)
        lvar, uvar, lcon, ucon = qp.lvar, qp.uvar, qp.lcon, qp.ucon
        if np.any(lvar == uvar):
            raise NotImplementedError("fixed variables (MadNLP.MakeParameter) are outside the hot-path scope")
        if not qp.minimize:
            raise NotImplementedError("maximization models (MadNLP obj_sign = -1) are outside the hot-path scope: negate c, H, c0")
        self.ind_ineq = np.flatnonzero(lcon != ucon)
        nx, ns = qp.nvar, len(self.ind_ineq)
        self.nx, self.ns = nx, ns
        self.n, self.m = nx + ns, qp.ncon
        n, m = self.n, self.m
        if opt.kkt_system == "Normal" and qp.nnzh > 0:
            raise ValueError("The KKT system NormalKKTSystem supports only linear programs.")  # normalkkt.jl:45-48
        ib = int(opt.index_base)
        self.A_V_host = np.concatenate([qp.Avals, -np.ones(ns)])
        # ---- Hessian operator (MadIPMOperator symmetric=true, cuda_wrapper.jl:62-68): host part
        if qp.nnzh > 0:
            hr, hc, hv = qp.Hrows.astype(np.int64), qp.Hcols.astype(np.int64), qp.Hvals
            off = hr != hc
            fr = np.concatenate([hr, hc[off]]).astype(np.int32)
            fc = np.concatenate([hc, hr[off]]).astype(np.int32)
            fv = np.concatenate([hv, hv[off]])
            Hp, Hj, Hmap = _lib.coo_to_csr(nx, nx, fr + ib, fc + ib, index_base=ib)
            self.H_full_host = fv[Hmap - ib]
        # Pinned staging of the numeric problem data on a helper thread while this thread does the host-only index work
        # below: page-locking ~100 MB takes tens of milliseconds, and the driver serialises it with device allocations,
        # so it is joined before the first one.
        import threading
        stage_err = []

        def _stage():
            try:
                torch.cuda.set_device(self.device)
                self._stage_problem()
            except Exception as exc:       # surfaced after the join
                stage_err.append(exc)
        stager = threading.Thread(target=_stage)
        stager.start()
        lfull = np.concatenate([lvar, lcon[self.ind_ineq]])
        ufull = np.concatenate([uvar, ucon[self.ind_ineq]])
        self.ind_lb = np.flatnonzero(np.isfinite(lfull)).astype(np.int64)
        self.ind_ub = np.flatnonzero(np.isfinite(ufull)).astype(np.int64)
        self.nlb, self.nub = len(self.ind_lb), len(self.ind_ub)
        nlb, nub = self.nlb, self.nub
        # ---- Jacobian with slack columns (normalkkt.jl:70-79) in CSR through coo_to_csr
        I = np.concatenate([qp.Arows, self.ind_ineq]).astype(np.int32)
        J = np.concatenate([qp.Acols, nx + np.arange(ns)]).astype(np.int32)
        self.A_I, self.A_J = I, J
        _mark("index sets")
        Ap, Aj, Amap = _lib.coo_to_csr(m, n, I + ib, J + ib, index_base=ib)      # in index_base like the inputs
        _mark("coo_to_csr")
        stager.join()
        if stage_err:
            raise stage_err[0]
        _mark("staging thread joined")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.h = _lib.Handle(device=opt.device, stream=stream)
        if opt.grid_limit:
            self.h.set_grid_limit(opt.grid_limit)
        dev = self.device
        z = lambda k: torch.zeros(max(k, 0), dtype=torch.float64, device=dev)
        # ---- iterate and buffers (structure.jl:125-153)
        self.x, self.xl, self.xu, self.zl, self.zu, self.f = z(n), z(n), z(n), z(n), z(n), z(n)
        self.y, self.c, self.rhs, self.jacl = z(m), z(m), z(m), z(n)
        N = n + m + nlb + nub
        self.d, self.p, self._w1, self._w2, self._w3 = z(N), z(N), z(N), z(N), z(N)
        self.correction_lb, self.correction_ub = z(nlb), z(nub)
        self.d_ind_lb = _dev(self.ind_lb + ib, dev, torch.int64)
        self.d_ind_ub = _dev(self.ind_ub + ib, dev, torch.int64)
        _mark("vectors")
        self.h.spmv_setup(m, n, Ap, Aj, index_base=ib)
        _mark("spmv_setup")
        self._Ap_abi, self._Aj_abi = Ap, Aj
        self.Ap, self.Aj, self.A_csr_map = Ap - ib, Aj - ib, Amap - ib            # 0-based copies for the host logic
        self.AT_x = z(len(Aj))                      # AT.nzVal: CSR-ordered values of A
        self.A_V = z(len(Aj))                       # A.V: COO-ordered values (jac + slack), like kkt.A.V
        self.d_A_csr_map = _dev(Amap, dev, torch.int64)                            # as returned (index_base)
        self.hH = None
        if qp.nnzh > 0:
            self.hH = self.h
            self.h.hess_setup(nx, Hp, Hj, index_base=ib)
            self.Hx = z(len(Hj))
        self.cvec = z(n)
        # ---- KKT system
        self.buffer_n, self.buffer_m = z(n), z(m)
        self.l_diag, self.u_diag, self.l_lower, self.u_lower = z(nlb), z(nub), z(nlb), z(nub)
        self.reg = z(n)
        if opt.kkt_system == "Normal":
            self.pr_diag, self.du_diag = z(n), z(m)
            _mark("buffers")
            Cp, Cj = self.h.normal_symbolic(m, n, index_base=ib)      # A as registered by spmv_setup above
            _mark("normal_symbolic")
            self.aug_colptr, self.aug_rowval = Cp, Cj
            if opt.linear_solver == "distributed":
                from .distributed import DistributedB200Solver
                if not (0 < opt.n_border <= m):
                    raise ValueError("linear_solver='distributed' needs n_border = number of linking rows (placed last)")
                self._aug_nz_ext = z(len(Cj) + 1)                 # one trailing zero slot for the rank-local gather
                self.aug_nz = self._aug_nz_ext[:len(Cj)]
                self.linear_solver = DistributedB200Solver(m, Cp - ib, Cj - ib, self._aug_nz_ext, opt.n_border, opt.device, stream)
            else:
                self.aug_nz = z(len(Cj))
                if opt.cudss_algorithm not in ("auto", "CHOLESKY", "LDL"):
                    raise ValueError("cudss_algorithm must be 'auto', 'CHOLESKY' or 'LDL'")
                kind = _lib.MIPM_LDL_DEFINITE if opt.cudss_algorithm == "LDL" else _lib.MIPM_CHOLESKY
                self.linear_solver = B200Solver(self.h, m, Cp, Cj, self.aug_nz, kind, opt.ordering, opt.ir_steps, ib)
        elif opt.kkt_system in ("K2", "K2.5"):
            # MadNLP.SparseKKTSystem (K2.5: ScaledSparseKKTSystem, same pattern): COO values [pr_diag; hess; jac(+slack);
            # du_diag], lower triangular
            if opt.cudss_algorithm not in ("auto", "LDL"):
                raise ValueError("the K2 / K2.5 systems are indefinite: cudss_algorithm must be 'auto' or 'LDL'")
            nnzh, nnzj = qp.nnzh, len(I)
            KI = np.concatenate([np.arange(n), qp.Hrows, n + I, n + np.arange(m)]).astype(np.int32)
            KJ = np.concatenate([np.arange(n), qp.Hcols, J, n + np.arange(m)]).astype(np.int32)
            colptr, rowval, kmap = self.h.k2_symbolic(n + m, KI + ib, KJ + ib, index_base=ib)
            self.aug_colptr, self.aug_rowval, self.aug_csc_map = colptr, rowval, kmap
            self.aug_raw_V = z(n + nnzh + nnzj + m)
            self.pr_diag = self.aug_raw_V[:n]                       # views, like MadNLP's _madnlp_unsafe_wrap
            self.hess = self.aug_raw_V[n:n + nnzh]
            self.jac = self.aug_raw_V[n + nnzh:n + nnzh + nnzj]
            self.du_diag = self.aug_raw_V[n + nnzh + nnzj:]
            self.aug_nz = z(len(rowval))
            self.linear_solver = B200Solver(self.h, n + m, colptr, rowval, self.aug_nz, _lib.MIPM_LDL, opt.ordering, opt.ir_steps, ib)
            if opt.kkt_system == "K2.5":
                # unscaled values + their index arrays on the device: build_kkt! scales them by S into the COO buffer
                self.scaling_factor = z(n)
                self.hess_raw, self.jac_raw = z(nnzh), z(nnzj)
                self.d_H_I, self.d_H_J = _dev(qp.Hrows + ib, dev, torch.int32), _dev(qp.Hcols + ib, dev, torch.int32)
                self.d_J_J = _dev(J + ib, dev, torch.int32)
        else:
            raise ValueError(opt.kkt_system)
        _mark("linear solver (analysis + device setup)")
        # ---- bind the device vectors once
        mv = MpcVectors()
        mv.n, mv.m, mv.nlb, mv.nub, mv.index_base = n, m, nlb, nub, ib
        P = lambda t: t.data_ptr() if t.numel() > 0 else None
        mv.d_ind_lb, mv.d_ind_ub = P(self.d_ind_lb), P(self.d_ind_ub)
        for name, t in [("d_x", self.x), ("d_xl", self.xl), ("d_xu", self.xu), ("d_zl", self.zl), ("d_zu", self.zu),
                        ("d_f", self.f), ("d_y", self.y), ("d_c", self.c), ("d_rhs", self.rhs), ("d_jacl", self.jacl),
                        ("d_d", self.d), ("d_p", self.p), ("d_w", self._w1), ("d_corr_lb", self.correction_lb),
                        ("d_corr_ub", self.correction_ub), ("d_reg", self.reg), ("d_pr_diag", self.pr_diag),
                        ("d_du_diag", self.du_diag), ("d_l_diag", self.l_diag), ("d_u_diag", self.u_diag),
                        ("d_l_lower", self.l_lower), ("d_u_lower", self.u_lower)]:
            setattr(mv, name, P(t))
        self._mv = mv
        self.h.mpc_bind(mv)
        md = MpcModel()
        md.kkt_kind = 0 if opt.kkt_system == "Normal" else 1        # (K2.5 runs the fine-grained sequence only)
        md.exact_order = int(opt.exact_assembly_order)
        md.nx, md.c0 = nx, 0.0
        md.d_ATx, md.d_cvec, md.d_aug_nz = P(self.AT_x), P(self.cvec), P(self.aug_nz)
        md.d_Hx = P(self.Hx) if qp.nnzh > 0 else None
        md.d_aug_raw_V = P(self.aug_raw_V) if opt.kkt_system in ("K2", "K2.5") else None
        md.d_buffer_n, md.d_buffer_m = P(self.buffer_n), P(self.buffer_m)
        self._md = md
        self.h.mpc_set_model(md)
        # a variable with no finite bound keeps pr_diag = del_w (1e-10): the KKT system is then badly
        # conditioned from the first iteration on, so the fused path refines every solve from the start
        _mark("bind")
        self._has_free = (nlb + nub > 0 or n > 0) and bool(np.any(~np.isfinite(lfull) & ~np.isfinite(ufull)))
        self._fused_started = False
        self._fused_ir = max(opt.ir_steps, 1 if self._has_free else 0)
        # ---- scalars (structure.jl:62-76)
        self.obj_val = 0.0
        self.inf_pr = self.inf_du = self.inf_compl = 0.0
        self.norm_b = self.norm_c = 0.0
        self.mu = self.mu_curr = 0.0
        self.alpha_p = self.alpha_d = 0.0
        self.del_w = self.del_c = 0.0
        self.best_complementarity = float("inf")
        self.status = INITIAL
        self.k = 0
        self.dnorm = 0.0
        self.obj_scale = 1.0
        self.con_scale = np.ones(m)
        self.trace = []
        self.cnt = dict(linear_solver_time=0.0, init_time=time.time() - t0, factorizations=0, solves=0)
        self._reg = self.opt.regularization

    # ------------------------------------------------------------------ model callbacks (device)
    def _eval_f(self):
        """obj = c0 + c'x + x'Hx/2 (MadIPMCUDAExt.jl:34-38)."""
        v = self.h.dot(self.nx, self.cvec, self.x)
        if self.hH is not None:
            self.hH.hess_spmv(1.0, self.Hx, self.x, 0.0, self.buffer_n)
            v += 0.5 * self.h.dot(self.nx, self.buffer_n, self.x)
        return self.obj_scale * self.qp.c0 + v

    def _eval_grad(self):
        """f = Hx + c (MadIPMCUDAExt.jl:40-45)."""
        self.h.copy(self.n, self.cvec, self.f)
        if self.hH is not None:
            self.hH.hess_spmv(1.0, self.Hx, self.x, 1.0, self.f)

    def _eval_cons(self):
        """c(x) = A x - s - rhs with slack columns inside A (App. A)."""
        self.h.copy(self.m, self.rhs, self.c)
        self.h.spmv(0, 1.0, self.AT_x, self.x, -1.0, self.c)

    def jtprod(self, out, y):
        """MadNLP.jtprod!(y, kkt, x) (normalkkt.jl:176-178)."""
        self.h.spmv(1, 1.0, self.AT_x, y, 0.0, out)

    # ------------------------------------------------------------------ KKT system
    def compress_jacobian(self):
        """normalkkt.jl:163-172 / cuda_wrapper.jl:32-41: AT.nzVal = A.V[A_csr_map] (+ slack = -1); A.V (the jac_coord!
        result, scaled by con_scale) was placed on the device by _madnlp_initialize."""
        self.h.gather(len(self.Aj), self.A_V, self.d_A_csr_map, self.AT_x, index_base=self.opt.index_base)     # AT.nzVal .= A.V[A_csr_map] (slack entries are -1)
        self.h.spmv_cache_values(self.AT_x)       # the values stay fixed until the next compress_jacobian!
        if self.opt.kkt_system == "Normal":
            self.h.normal_set_jacobian(self.AT_x)
        elif self.opt.kkt_system == "K2.5":
            self.h.copy(len(self.Aj), self.A_V, self.jac_raw)
        else:
            self.h.copy(len(self.Aj), self.A_V, self.jac)

    def compress_hessian(self):
        """hess_coord! values times obj_scale (the raw values were uploaded by _madnlp_initialize)."""
        if self.qp.nnzh > 0 and self.obj_scale != 1.0:
            self.h.axpby(self.Hx.numel(), self.obj_scale, self.Hx, 0.0, self.Hx)
            if self.opt.kkt_system in ("K2", "K2.5"):
                hv = self.hess if self.opt.kkt_system == "K2" else self.hess_raw
                self.h.axpby(hv.numel(), self.obj_scale, hv, 0.0, hv)

    def build_kkt(self):
        """build_kkt!: normalkkt.jl:180-194 (Normal) / MadNLP.transfer! (K2)."""
        if self.opt.kkt_system == "Normal":
            self.h.normal_assemble(self.pr_diag, self.aug_nz, self.opt.exact_assembly_order)
        else:
            if self.opt.kkt_system == "K2.5":       # S W S and J S into the COO buffer, then the same transfer
                self.h.k25_scale_values(self.d_H_I, self.d_H_J, self.hess_raw, self.hess, self.d_J_J, self.jac_raw, self.jac,
                                        self.scaling_factor, index_base=self.opt.index_base)
            self.h.k2_transfer(self.aug_raw_V, self.aug_nz)

    def factorize_wrapper(self):
        """MadNLP.factorize_wrapper!: build_kkt! + factorize!."""
        self.build_kkt()
        self.linear_solver.factorize()
        self.cnt["factorizations"] += 1

    def kkt_solve(self, w):
        """solve!(kkt, w): normalkkt.jl:196-219 (Normal) / MadNLP (K2)."""
        h = self.h
        if self.opt.kkt_system == "Normal":
            h.normal_solve_stage(0, w, self.buffer_n, self.buffer_m)
            h.spmv(0, 1.0, self.AT_x, self.buffer_n, -1.0, self.buffer_m)     # A Sigma^-1 r1 - r2
            self.linear_solver.solve(self.buffer_m)
            h.normal_solve_stage(1, w, self.buffer_n, self.buffer_m)
            h.spmv(1, -1.0, self.AT_x, self.buffer_m, 1.0, self.buffer_n)     # r1 - A' dy
            h.normal_solve_stage(2, w, self.buffer_n, self.buffer_m)
        elif self.opt.kkt_system == "K2.5":
            h.reduce_rhs_scaled(w, self.scaling_factor)
            self.linear_solver.solve(w)
            h.finish_aug_solve_scaled(w, self.scaling_factor)
        else:
            h.reduce_rhs(w)
            self.linear_solver.solve(w)       # primal_dual(w) = first n+m entries, in place
            h.finish_aug_solve(w)
        self.cnt["solves"] += 1
        return w

    def kkt_mul(self, w, v, alpha, beta):
        """mul!(w, kkt, v, alpha, beta): normalkkt.jl:221-233 (+ Hessian block for K2)."""
        n, m = self.n, self.m
        h = self.h
        h.spmv_pair(self.AT_x, alpha, v[:n], beta, w[n:n + m], alpha, v[n:n + m], beta, w[:n])     # A v_x and A' v_y in one launch
        if self.hH is not None and self.opt.kkt_system in ("K2", "K2.5"):
            self.hH.hess_spmv(alpha, self.Hx, v[:self.nx], 1.0, w[:self.nx])
        if self.opt.kkt_system == "K2.5":
            h.kktmul_scaled(w, v, alpha, beta)
        else:
            h.kktmul(w, v, alpha, beta)
        return w

    # ------------------------------------------------------------------ src/linear_solver.jl
    def solve_system(self):
        """src/linear_solver.jl:19-44."""
        N = self.d.numel()
        self.h.copy(N, self.p, self.d)
        self.kkt_solve(self.d)
        self.h.copy(N, self.p, self._w1)
        self.kkt_mul(self._w1, self.d, -1.0, 1.0)
        norm_w, norm_p = self.h.residual_norms(self._w1, self.p)
        self.residual_ratio = norm_w / max(1.0, norm_p)
        nref = 0
        while self.residual_ratio > self.opt.refine_tol and nref < self.opt.max_refine:
            # d += K^-1 (p - K d), then the reference's residual check again. A round that does not reduce the residual is
            # taken back: on ill-conditioned systems the correction carries the same error as the solve it corrects.
            self.h.copy(N, self._w1, self._w3)
            self.kkt_solve(self._w3)
            self.h.axpby(N, 1.0, self._w3, 1.0, self.d)
            self.h.copy(N, self.p, self._w1)
            self.kkt_mul(self._w1, self.d, -1.0, 1.0)
            norm_w, norm_p = self.h.residual_norms(self._w1, self.p)
            prev, self.residual_ratio = self.residual_ratio, norm_w / max(1.0, norm_p)
            nref += 1
            self.cnt["refinements"] = self.cnt.get("refinements", 0) + 1
            if not (self.residual_ratio < prev):
                self.h.axpby(N, -1.0, self._w3, 1.0, self.d)
                self.residual_ratio = prev
                self.cnt["refinements_rejected"] = self.cnt.get("refinements_rejected", 0) + 1
                break
            if not (self.residual_ratio < 0.5 * prev):
                break
        if np.isnan(self.residual_ratio) or (self.opt.check_residual and self.residual_ratio > self.opt.tol_linear_solve):
            raise SolveException("residual %.3e" % self.residual_ratio)
        return self.d

    def factorize_regularized_system(self):
        """src/linear_solver.jl:6-17."""
        t0 = time.perf_counter()
        for _ in range(3):
            if self.opt.kkt_system == "K2.5":
                self.h.set_aug_diagonal_reg_scaled(self.del_w, self.del_c, self.scaling_factor)
            else:
                self.h.set_aug_diagonal_reg(self.del_w, self.del_c)
            self.factorize_wrapper()
            if self.linear_solver.is_factorized():
                break
            self.del_w *= 100.0
            self.del_c *= 100.0
        self.cnt["linear_solver_time"] += time.perf_counter() - t0

    # ------------------------------------------------------------------ regularization (kernels.jl:364-401)
    def init_regularization(self):
        self.del_w = 1.0
        self.del_c = 0.0 if isinstance(self._reg, NoRegularization) else self._reg.delta_d

    def update_regularization(self):
        r = self._reg
        if isinstance(r, NoRegularization):
            self.del_w, self.del_c = 0.0, 0.0
        elif isinstance(r, FixedRegularization):
            self.del_w, self.del_c = r.delta_p, r.delta_d
        elif isinstance(r, AdaptiveRegularization):
            r.delta_p = max(r.delta_p / 10.0, r.delta_min)
            r.delta_d = min(r.delta_d / 10.0, -r.delta_min)
            self.del_w, self.del_c = r.delta_p, r.delta_d
        else:
            raise TypeError(r)

    # ------------------------------------------------------------------ initialization
    def _stage_problem(self):
        """Pinned host copies of the numeric problem data (the host side of `convert(QuadraticModel{T, CuVector}, qp)`,
        README.md:77): every solve() uploads them again, so a timed solve includes its host->device traffic."""
        qp, nx, ns, m = self.qp, self.nx, self.ns, self.m
        # ONE page-locked slab for all arrays (a pinned allocation per array costs tens of milliseconds each); small
        # problems stay in pageable memory (page-locking costs more than their copies)
        zs = np.zeros(ns)
        arrays = dict(
            x0=np.concatenate([qp.x0, zs]), c=np.concatenate([qp.c, zs]), y0=qp.y0,
            xl=np.concatenate([qp.lvar, qp.lcon[self.ind_ineq]]), xu=np.concatenate([qp.uvar, qp.ucon[self.ind_ineq]]),
            rhs=np.where(qp.lcon == qp.ucon, qp.lcon, 0.0), A_V=self.A_V_host)
        if qp.nnzh > 0:
            arrays["H_full"] = self.H_full_host
            arrays["H_tril"] = qp.Hvals
        total = sum(len(a) for a in arrays.values())
        slab = torch.empty(max(total, 1), dtype=torch.float64, pin_memory=(total >= 65536))
        self._host, off = {}, 0
        for name, a in arrays.items():
            view = slab[off:off + len(a)]
            view.numpy()[:] = a
            self._host[name] = view
            off += len(a)
        self._amax_A = float(np.abs(self.A_V_host).max()) if len(self.A_V_host) else 0.0
        self.h2d_bytes_per_solve = int(sum(t.numel() * 8 for t in self._host.values()))
        # page-locked landing zone of the results (x, constraints, y, zl, zu), two of them used in turn: the arrays of an
        # ExecutionStats stay valid until the solve after next on the same solver
        rtotal = 3 * (nx + ns) + 2 * m
        self._res_slabs = [torch.empty(max(rtotal, 1), dtype=torch.float64, pin_memory=(rtotal >= 65536)) for _ in range(2)]
        self._res_turn = 0

    def _madnlp_initialize(self):
        """MadNLP.initialize!(cb, ...) + set_scaling! (App. B) on the device; inputs come from pinned host memory."""
        qp, opt, nx, n, m, h = self.qp, self.opt, self.nx, self.n, self.m, self.h
        H = self._host
        up_ = lambda t, a: t.copy_(a, non_blocking=True)
        up_(self.x, H["x0"]), up_(self.xl, H["xl"]), up_(self.xu, H["xu"]), up_(self.rhs, H["rhs"]), up_(self.cvec, H["c"])
        up_(self.y, H["y0"]), up_(self.A_V, H["A_V"])
        if qp.nnzh > 0:
            up_(self.Hx, H["H_full"])
            if opt.kkt_system == "K2":
                up_(self.hess, H["H_tril"])
            elif opt.kkt_system == "K2.5":
                up_(self.hess_raw, H["H_tril"])
        h.init_bounds(n, opt.bound_relax_factor, opt.bound_push, opt.bound_fac, self.x, self.xl, self.xu)
        self.con_scale = np.ones(m)
        self.obj_scale = 1.0
        if opt.scaling:
            # con_scale_i = min(1, 100 / max_j |A_ij|) is identically 1 when max |A_ij| <= 100 (the usual case, decided
            # once from the values); only otherwise are the per-row maxima needed (host, CSR order)
            if self._amax_A > 100.0:
                absv = np.abs(self.A_V_host)[self.A_csr_map]
                rowmax = np.zeros(m)
                nz = np.flatnonzero(np.diff(self.Ap) > 0)
                rowmax[nz] = np.maximum.reduceat(absv, self.Ap[:-1][nz])
                self.con_scale = np.minimum(1.0, 100.0 / np.maximum(rowmax, 1e-300))
                # MadNLP.set_scaling! (MadNLP 0.8): rhs .*= con_scale; y0 ./= con_scale; slack(x), slack(xl), slack(xu)
                # .*= con_scale[ind_ineq]; the Jacobian entries of the model are scaled by row, the slack entries stay
                # -1 (compress_jacobian!, normalkkt.jl:163-172). Rare path (some |A_ij| > 100): host arithmetic.
                self.rhs.copy_(torch.from_numpy(H["rhs"].numpy() * self.con_scale))
                self.y.copy_(torch.from_numpy(H["y0"].numpy() / self.con_scale))
                csv = np.concatenate([self.con_scale[self.A_I[:qp.nnzj]], np.ones(self.ns)])
                self.A_V.copy_(torch.from_numpy(self.A_V_host * csv))
                if self.ns:
                    cs = torch.from_numpy(self.con_scale[self.ind_ineq]).to(self.device)
                    self.x[nx:] *= cs
                    self.xl[nx:] *= cs
                    self.xu[nx:] *= cs
            # obj_scale = min(1, 100 / ||grad f(x0)||_inf) with grad f = c + H x0
            if qp.nnzh > 0:
                h.copy(n, self.cvec, self.f)
                h.hess_spmv(1.0, self.Hx, self.x, 1.0, self.f)
                gn = h.amax(n, self.f)
            else:
                gn = h.amax(n, self.cvec)
            self.obj_scale = min(1.0, 100.0 / gn) if gn > 0 else 1.0
        self._unit_con_scale = bool(np.all(self.con_scale == 1.0)) if self._amax_A > 100.0 else True
        if self.obj_scale != 1.0:
            h.axpby(n, self.obj_scale, self.cvec, 0.0, self.cvec)
        self.zl.zero_(), self.zu.zero_()
        self.norm_b = h.amax(m, self.rhs)

    def initialize(self):
        """src/solver.jl:127-189."""
        opt, h = self.opt, self.h
        self._madnlp_initialize()
        h.fill(self.n, 0.0, self.jacl)
        # MadNLP.initialize!(kkt) (normalkkt.jl:150-161)
        h.fill(self.n, 1.0, self.reg), h.fill(self.n, 1.0, self.pr_diag), h.fill(self.m, 0.0, self.du_diag)
        h.fill(self.nlb, 0.0, self.l_lower), h.fill(self.nub, 0.0, self.u_lower)
        h.fill(self.nlb, 1.0, self.l_diag), h.fill(self.nub, 1.0, self.u_diag)
        if opt.kkt_system == "K2.5":
            h.fill(self.n, 1.0, self.scaling_factor)
        self.init_regularization()
        self.compress_hessian()
        self.compress_jacobian()
        self.obj_val = self._eval_f()
        self._eval_grad()
        self._eval_cons()
        # norm_c = ||grad f(x0)||_inf (quirk A.9 x)
        self.norm_c = float(self.f.abs().max().item()) if self.n else 0.0   # one-time, init only
        self.init_starting_point()
        self.mu = opt.mu_init
        self.best_complementarity = float("inf")
        self.status = REGULAR
        self.jtprod(self.jacl, self.y)
        self._fused_started = False
        self._fused_ir = max(self.opt.ir_steps, 1 if self._has_free else 0)

    def init_starting_point(self):
        """src/solver.jl:6-125."""
        h, n, m = self.h, self.n, self.m
        N = self.p.numel()
        h.fill(n, self.del_w, self.reg), h.fill(n, self.del_w, self.pr_diag), h.fill(m, self.del_c, self.du_diag)
        self.factorize_wrapper()
        if not self.linear_solver.is_factorized():
            raise SolveException("initial factorization failed")
        # Step 1 (kernels.jl:1-9): p = [0; -c; 0; 0]
        h.fill(N, 0.0, self.p)
        h.axpby(m, -1.0, self.c, 0.0, self.p[n:n + m])
        self.solve_system()
        h.axpby(n, 1.0, self.d[:n], 1.0, self.x)
        # Step 2 (kernels.jl:11-19): p = [-f; 0; 0; 0]
        h.fill(N, 0.0, self.p)
        h.axpby(n, -1.0, self.f, 0.0, self.p[:n])
        self.solve_system()
        h.copy(m, self.d[n:n + m], self.y)
        # Step 3: res = A'y + f, held in jacl like the reference (solver.jl:13,37-39)
        self.jtprod(self.jacl, self.y)
        h.axpby(n, 1.0, self.f, 1.0, self.jacl)
        mins = h.init_point_stage(0)
        delta_x = max(0.0, -1.5 * mins[0], -1.5 * mins[1])
        delta_s = max(0.0, -1.5 * mins[2], -1.5 * mins[3])
        s = h.init_point_stage(1, delta_x, delta_s)
        mu = s[0]
        with np.errstate(divide="ignore", invalid="ignore"):
            delta_x2 = float(np.float64(mu) / (2 * (np.float64(s[1]) + s[2])))
            delta_s2 = float(np.float64(mu) / (2 * (np.float64(s[3]) + s[4])))
        chk = h.init_point_stage(2, delta_x2, delta_s2, self.opt.bound_fac)
        if self.nlb > 0 and not (chk[0] > 0.0 and chk[2] > 0.0):
            raise AssertionError("starting point not strictly interior (lower)")
        if self.nub > 0 and not (chk[1] > 0.0 and chk[3] > 0.0):
            raise AssertionError("starting point not strictly interior (upper)")

    # ------------------------------------------------------------------ MPC (src/solver.jl:194-360)
    def update_termination_criteria(self):
        opt = self.opt
        dobj, nc, ndu, ncompl, dnorm = self.h.termination_measures()
        self.dobj = dobj
        self.dnorm = dnorm
        self.inf_pr = nc / max(1.0, self.norm_b)
        self.inf_du = ndu / max(1.0, self.norm_c)
        self.inf_compl = ncompl / max(1.0, self.norm_c)
        self.best_complementarity = min(self.best_complementarity, self.inf_compl)
        if max(self.inf_pr, self.inf_du, self.inf_compl) <= opt.tol:
            self.status = SOLVE_SUCCEEDED
        elif (self.inf_compl > opt.divergence_tol * self.best_complementarity) and (dobj > max(10.0 * abs(self.obj_val), 1.0)):
            self.status = INFEASIBLE_PROBLEM_DETECTED
        elif self.obj_val < -opt.divergence_tol * max(10.0, abs(dobj), 1.0):
            self.status = DIVERGING_ITERATES
        elif self.k >= opt.max_iter:
            self.status = MAXIMUM_ITERATIONS_EXCEEDED
        elif time.time() - self.start_time >= opt.max_wall_time:
            self.status = MAXIMUM_WALLTIME_EXCEEDED

    def get_fraction_to_boundary_step(self, tau):
        """src/kernels.jl:274-289."""
        a, _ = self.h.get_alpha_max(tau)
        return min(a[0], a[1]), min(a[2], a[3])

    def update_barrier(self, mu_affine):
        """src/kernels.jl:210-220 (quirk A.9 vii: has_inequalities == nlb+nub > 0)."""
        mu_curr = self.h.get_complementarity_measure()
        if self.nlb + self.nub > 0:
            sigma = min(max((mu_affine / mu_curr) ** 3, 1e-6), 10.0)
        else:
            sigma = 1.0
        self.mu = max(self.opt.mu_min, sigma * mu_curr)
        return mu_curr

    def prediction_step(self):
        """src/solver.jl:230-237."""
        self.h.set_predictive_rhs()
        self.solve_system()
        ap, ad = self.get_fraction_to_boundary_step(1.0)
        mu_affine = self.h.get_affine_complementarity_measure(ap, ad)
        self.h.get_correction()
        self.mu_curr = self.update_barrier(mu_affine)

    def mehrotra_correction_direction(self):
        """src/solver.jl:239-243."""
        self.h.set_correction_rhs(self.mu)
        self.solve_system()

    def gondzio_correction_direction(self):
        """src/solver.jl:245-298."""
        if self.opt.max_ncorr <= 0:
            return
        delta, bmin, bmax, tau = 0.1, 0.1, 10.0, 0.995
        N = self.d.numel()
        alpha_p, alpha_d = self.get_fraction_to_boundary_step(tau)
        for _ in range(self.opt.max_ncorr):
            tap, tad = min(alpha_p + delta, 1.0), min(alpha_d + delta, 1.0)
            ga = self.h.get_affine_complementarity_measure(tap, tad)
            g = self.mu_curr
            mu = (ga / g) ** 2 * ga
            self.h.set_extra_correction(tap, tad, bmin, bmax, mu)
            self.h.set_correction_rhs(mu)
            self.h.copy(N, self.d, self._w2)
            self.solve_system()
            hap, had = self.get_fraction_to_boundary_step(tau)
            if (hap < 1.005 * alpha_p) or (had < 1.005 * alpha_d):
                self.h.copy(N, self._w2, self.d)
                break
            alpha_p, alpha_d = hap, had

    def update_step_size(self):
        """src/kernels.jl:291-358."""
        rule = self.opt.step_rule
        if isinstance(rule, ConservativeStep):
            self.alpha_p, self.alpha_d = self.get_fraction_to_boundary_step(rule.tau)
        elif isinstance(rule, AdaptiveStep):
            tau = max(1 - self.mu, rule.tau_min)
            self.alpha_p, self.alpha_d = self.get_fraction_to_boundary_step(tau)
        elif isinstance(rule, MehrotraAdaptiveStep):
            self._mehrotra_adaptive_step(rule)
        else:
            raise TypeError(rule)

    def _mehrotra_adaptive_step(self, rule):
        """src/kernels.jl:309-358 on the device (mipm_mehrotra_adaptive_step): the reference's scalar indexing into device
        arrays from the host is done by one device thread."""
        self.alpha_p, self.alpha_d = self.h.mehrotra_adaptive_step(rule.gamma_f)

    def apply_step(self):
        """src/solver.jl:308-317."""
        self.h.apply_step(self.alpha_p, self.alpha_d, self.mu)
        self.k += 1

    def evaluate_model(self):
        """src/solver.jl:319-326."""
        self.obj_val = self._eval_f()
        self._eval_cons()
        self._eval_grad()
        self.jtprod(self.jacl, self.y)

    def _record(self):
        self.trace.append(dict(
            k=self.k, objective=self.obj_val / self.obj_scale, dual_objective=self.dobj / self.obj_scale,
            inf_pr=self.inf_pr, inf_du=self.inf_du, inf_compl=self.inf_compl, mu=self.mu,
            alpha_p=self.alpha_p, alpha_d=self.alpha_d, del_w=self.del_w,
            dnorm=0.0 if self.k == 0 else self.dnorm))

    def _use_fused(self):
        return (self.opt.fused and self.opt.max_ncorr <= 0 and not self.opt.check_residual
                and self.opt.kkt_system != "K2.5"
                and isinstance(self.opt.step_rule, (AdaptiveStep, ConservativeStep, MehrotraAdaptiveStep)))

    def _mpc_iteration_fused(self):
        """The same loop body through mipm_mpc_iter_begin / mipm_mpc_iter_rest: one host sync."""
        opt = self.opt
        trace_del_w = self.del_w
        self.update_regularization()                      # host-side schedule (kernels.jl:370-401)
        ext = opt.linear_solver != "b200"                 # external (distributed) linear solver: phased entry points
        # Close to convergence (previous measures within 10 x tol) the termination measures are read before the next KKT
        # system is assembled and factorized, so the iteration that detects convergence does not pay for a factorization
        # it never uses; otherwise both go out together and the iteration has one synchronisation.
        peek = (self._fused_started and self.status == REGULAR and
                max(self.inf_pr, self.inf_du, self.inf_compl) <= 10.0 * opt.tol)
        ok = True
        if peek:
            out = self.h.mpc_peek()
        elif ext:
            self.h.mpc_ext_begin(self.del_w, self.del_c)
            self.linear_solver.factorize()
            ok = self.linear_solver.is_factorized()
            out = self.h.mpc_ext_fetch()
        else:
            out, ok = self.h.mpc_iter_begin(self.del_w, self.del_c)
        if self._fused_started:                            # scalars of the step taken in the previous call
            self.obj_val = self.obj_scale * self.qp.c0 + out[5] + 0.5 * out[6]
            self.alpha_p, self.alpha_d, self.mu, self.mu_curr = out[7], out[8], out[9], out[10]
            for nw, npp in ((out[11], out[12]), (out[13], out[14])):
                ratio = nw / max(1.0, npp)
                if np.isnan(ratio):
                    raise SolveException("NaN residual after linear solve")
                if ratio > opt.refine_tol:
                    self._fused_ir = max(self._fused_ir, 1)   # later solves refine on the reduced system
            self.residual_ratio = ratio
        dobj, nc, ndu, ncompl, dnorm = out[:5]
        self.dobj, self.dnorm = dobj, dnorm
        self.inf_pr = nc / max(1.0, self.norm_b)
        self.inf_du = ndu / max(1.0, self.norm_c)
        self.inf_compl = ncompl / max(1.0, self.norm_c)
        self.best_complementarity = min(self.best_complementarity, self.inf_compl)
        if max(self.inf_pr, self.inf_du, self.inf_compl) <= opt.tol:
            self.status = SOLVE_SUCCEEDED
        elif (self.inf_compl > opt.divergence_tol * self.best_complementarity) and (dobj > max(10.0 * abs(self.obj_val), 1.0)):
            self.status = INFEASIBLE_PROBLEM_DETECTED
        elif self.obj_val < -opt.divergence_tol * max(10.0, abs(dobj), 1.0):
            self.status = DIVERGING_ITERATES
        elif self.k >= opt.max_iter:
            self.status = MAXIMUM_ITERATIONS_EXCEEDED
        elif time.time() - self.start_time >= opt.max_wall_time:
            self.status = MAXIMUM_WALLTIME_EXCEEDED
        new_del_w, self.del_w = self.del_w, trace_del_w
        self._record()
        self.del_w = new_del_w
        if self.status != REGULAR:
            return False
        if peek and ext:
            self.h.mpc_ext_begin(self.del_w, self.del_c)
            self.linear_solver.factorize()
            ok = self.linear_solver.is_factorized()
        elif peek:
            _, ok = self.h.mpc_iter_begin(self.del_w, self.del_c)
        self.cnt["factorizations"] += 1
        for _ in range(2):                                 # factorize_regularized_system! retries (linear_solver.jl:6-17)
            if ok:
                break
            self.del_w *= 100.0
            self.del_c *= 100.0
            if ext:
                self.h.set_aug_diagonal_reg(self.del_w, self.del_c)
                self.factorize_wrapper()
                ok = self.linear_solver.is_factorized()
            else:
                ok = self.h.mpc_refactor(self.del_w, self.del_c)
            self.cnt["factorizations"] += 1
        rule = opt.step_rule
        code, par = ((0, rule.tau_min) if isinstance(rule, AdaptiveStep) else
                     (2, rule.gamma_f) if isinstance(rule, MehrotraAdaptiveStep) else (1, rule.tau))
        if ext:
            self.h.mpc_ext_phase(0, opt.mu_min, code, par)
            self.linear_solver.solve(self.buffer_m)
            self.h.mpc_ext_phase(1, opt.mu_min, code, par)
            self.linear_solver.solve(self.buffer_m)
            self.h.mpc_ext_phase(2, opt.mu_min, code, par)
        else:
            self.h.mpc_iter_rest(opt.mu_min, code, par, self._fused_ir)
        self.cnt["solves"] += 2
        self._fused_started = True
        self.k += 1
        return True

    def mpc_iteration(self):
        """One pass of the loop body of mpc! (src/solver.jl:333-359). Returns False when done."""
        if self._use_fused():
            return self._mpc_iteration_fused()
        self.update_termination_criteria()
        self._record()
        if self.status != REGULAR:
            return False
        self.update_regularization()
        self.factorize_regularized_system()
        self.prediction_step()
        self.mehrotra_correction_direction()
        self.gondzio_correction_direction()
        self.update_step_size()
        self.apply_step()
        self.evaluate_model()
        return True

    def mpc(self):
        """src/solver.jl:332-360."""
        while self.mpc_iteration():
            pass

    def solve(self):
        """solve!(solver): src/solver.jl:362-418. The arrays of the returned ExecutionStats live in page-locked buffers
        owned by the solver (two sets, used in turn): they stay valid until the solve after next on this solver."""
        self.start_time = time.time()
        t0 = time.perf_counter()
        try:
            self.initialize()
            torch.cuda.synchronize(self.device)
            self.cnt["initialize_time"] = time.perf_counter() - t0
            self.mpc()
        except SolveException:
            # the reference throws the exception *type*, so its LinearSolverException branch is
            # missed and the status becomes INTERNAL_ERROR (solver.jl:396-404)
            self.status = INTERNAL_ERROR
            if self.opt.rethrow_error:
                raise
        except AssertionError:
            self.status = INTERNAL_ERROR
            if self.opt.rethrow_error:
                raise
        torch.cuda.synchronize(self.device)
        total = time.perf_counter() - t0
        # stats.constraints = A0 x (unscaled, without the slack columns): device SpMV, then undo both
        self.h.spmv(0, 1.0, self.AT_x, self.x, 0.0, self.buffer_m)
        # results: device -> page-locked slab in one go (no pageable staging, no fresh host pages), post-processed in place
        n, m = self.n, self.m
        self._res_turn ^= 1
        slab = self._res_slabs[self._res_turn]
        parts, off = [], 0
        for src, ln in ((self.x, n), (self.buffer_m, m), (self.y, m), (self.zl, n), (self.zu, n)):
            dst = slab[off:off + ln]
            if ln:
                dst.copy_(src, non_blocking=True)
            parts.append(dst.numpy())
            off += ln
        torch.cuda.synchronize(self.device)
        x, cons, mult, zl, zu = parts
        if self.ns:
            cons[self.ind_ineq] += x[self.nx:]         # the (scaled) slack columns are -1
        if not getattr(self, "_unit_con_scale", True):
            cons /= self.con_scale
            mult *= self.con_scale
        if self.obj_scale != 1.0:
            mult /= self.obj_scale
            zl /= self.obj_scale
            zu /= self.obj_scale
        return ExecutionStats(
            status=self.status, iter=self.k, objective=self.obj_val / self.obj_scale,
            dual_objective=getattr(self, "dobj", float("nan")) / self.obj_scale,
            solution=x[: self.nx], constraints=cons,
            multipliers=mult,
            multipliers_L=zl[: self.nx],
            multipliers_U=zu[: self.nx],
            trace=self.trace, total_time=total, linear_solver_time=self.cnt["linear_solver_time"],
            counters=dict(self.cnt, launches=self.h.launch_count(), ls_stats=self.linear_solver.stats),
        )


def solve(solver, **kwargs):
    """MadIPM.solve!(solver; kwargs...): runtime option overrides then solve (solver.jl:370-372)."""
    for k, v in kwargs.items():
        setattr(solver.opt, k, v)
    return solver.solve()


def madipm(qp, **kwargs):
    """madipm(m; kwargs...): src/solver.jl:425-428."""
    return MPCSolver(qp, **kwargs).solve()
