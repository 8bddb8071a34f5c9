"""oracle/mpc_oracle.py -- TEST INFRASTRUCTURE ONLY (CPU, numpy/scipy; never on the product path).

CPU restatement of MadIPM's Mehrotra predictor-corrector loop, written from the reference's
Julia source (klamike/MadIPM.jl at /root/reference) function by function. Every method cites
the file:line it follows. MadNLP 0.8.12 (Project.toml:30) is NOT vendored under
/root/reference; its pieces used here (initialize!, set_scaling!, get_index_constraints,
reduce_rhs!, finish_aug_solve!, _kktmul!, get_inf_*, adjust_boundary!, SparseKKTSystem's COO
layout) are restated from SURVEY.md Appendix A/B, i.e. from the algebra of the unreduced
Newton system, and are cross-checked in tests by K*d == p residuals.

PARITY STATUS: pinned only by the reference's one self-contained fixture, `simple_lp`
(test/runtests.jl:29-60: objective 1.0; Normal == K2 at 1e-6, :182-197). Everything else is
"parity unpinned" against the Julia reference (no Julia runtime in this image) and is instead
checked against scipy HiGHS final objectives and algebraic identities.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import time
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import sparse_ref

# MadNLP.Status subset (values are ours; names are MadNLP's)
REGULAR = "REGULAR"
SOLVE_SUCCEEDED = "SOLVE_SUCCEEDED"
INFEASIBLE_PROBLEM_DETECTED = "INFEASIBLE_PROBLEM_DETECTED"
DIVERGING_ITERATES = "DIVERGING_ITERATES"
MAXIMUM_ITERATIONS_EXCEEDED = "MAXIMUM_ITERATIONS_EXCEEDED"
MAXIMUM_WALLTIME_EXCEEDED = "MAXIMUM_WALLTIME_EXCEEDED"
INTERNAL_ERROR = "INTERNAL_ERROR"


@dataclass
class Options:
    """src/utils.jl:69-105 (IPMOptions defaults)."""
    tol: float = 1e-8
    kkt_system: str = "K2"              # "K2" = MadNLP.SparseKKTSystem (default), "Normal" = NormalKKTSystem
    # "ldl" = Davis LDL^T (the LDLSolver analogue, no pivoting) | "splu" = scipy SuperLU (pivoting,
    # the analogue of the reference's default CPU solver MUMPS) | "auto" = ldl for Normal, splu for K2
    # (static-pivot LDL^T is unsafe on K2 with the reference's default delta_c = +1e-10, quirk A.9 v)
    linear_solver: str = "auto"
    max_iter: int = 3000
    max_wall_time: float = 1e6
    divergence_tol: float = 1e4
    scaling: bool = True
    bound_push: float = 1e-2
    bound_fac: float = 1e-2
    bound_relax_factor: float = 1e-12
    regularization: tuple = ("fixed", 1e-10, 1e-10)   # FixedRegularization(1e-10, 1e-10), utils.jl:91
    step_rule: tuple = ("adaptive", 0.99)             # AdaptiveStep(0.99), utils.jl:93
    max_ncorr: int = 0
    mu_init: float = 1e-1
    mu_min: float = 1e-12
    tol_linear_solve: float = 1e-8
    check_residual: bool = False
    ordering: str = "rcm"               # fill-reducing ordering for "ldl": "rcm" | "natural"
    # False: literal O(m^2) build_normal_system (src/utils.jl:209-274). True: same canonical pattern
    # from a sparse product (needed at m >= 1e5 where the literal scan takes hours; equality of
    # the two is tested at small sizes)
    fast_symbolic: bool = False


@dataclass
class Stats:
    status: str
    iter: int
    objective: float
    dual_objective: float
    solution: np.ndarray
    constraints: np.ndarray
    multipliers: np.ndarray
    multipliers_L: np.ndarray
    multipliers_U: np.ndarray
    trace: list = field(default_factory=list)
    total_time: float = 0.0
    linear_solver_time: float = 0.0
    timers: dict = field(default_factory=dict)


class _LinearSolver:
    """MadNLP.AbstractLinearSolver mirror on a lower-triangular CSC."""

    def __init__(self, n, colptr, rowval, kind, ordering):
        self.n, self.colptr, self.rowval, self.kind = n, colptr, rowval, kind
        self.ok = False
        if kind == "ldl":
            perm = None
            if ordering == "rcm" and n > 2:
                from scipy.sparse.csgraph import reverse_cuthill_mckee
                pat = sp.csc_matrix((np.ones(len(rowval)), rowval, colptr), shape=(n, n))
                perm = reverse_cuthill_mckee((pat + pat.T).tocsr(), symmetric_mode=True).astype(np.int32)
            self.ldl = sparse_ref.LDL(n, colptr, rowval, perm)
            self.nnzL, self.flops = self.ldl.nnzL, self.ldl.flops

    def factorize(self, nzval):
        if self.kind == "ldl":
            self.ok = self.ldl.factorize(nzval)
        else:
            low = sp.csc_matrix((nzval, self.rowval, self.colptr), shape=(self.n, self.n))
            full = low + sp.tril(low, -1).T
            try:
                self.lu = spla.splu(full.tocsc())
                self.ok = True
            except RuntimeError:
                self.ok = False
        return self.ok

    def solve(self, b):
        return self.ldl.solve(b) if self.kind == "ldl" else self.lu.solve(b)


def get_index_constraints(lvar, uvar, lcon, ucon):
    """MadNLP.get_index_constraints (App. B): index sets over [x; s]."""
    ind_ineq = np.flatnonzero(lcon != ucon)
    lfull = np.concatenate([lvar, lcon[ind_ineq]])
    ufull = np.concatenate([uvar, ucon[ind_ineq]])
    ind_lb = np.flatnonzero(np.isfinite(lfull))
    ind_ub = np.flatnonzero(np.isfinite(ufull))
    ind_llb = np.flatnonzero(np.isfinite(lfull) & ~np.isfinite(ufull))
    ind_uub = np.flatnonzero(~np.isfinite(lfull) & np.isfinite(ufull))
    ind_fixed = np.flatnonzero(lvar == uvar)
    return dict(ind_ineq=ind_ineq, ind_lb=ind_lb, ind_ub=ind_ub, ind_llb=ind_llb,
                ind_uub=ind_uub, ind_fixed=ind_fixed)


class MPCOracle:
    """MPCSolver + solve! restated (src/structure.jl:79-178, src/solver.jl)."""

    def __init__(self, qp, **kwargs):
        self.opt = Options(**kwargs)
        self.qp = qp
        ic = get_index_constraints(qp.lvar, qp.uvar, qp.lcon, qp.ucon)
        if len(ic["ind_fixed"]) > 0:
            raise NotImplementedError("fixed variables (MadNLP.MakeParameter) are outside the hot-path scope")
        self.ind_ineq, self.ind_lb, self.ind_ub = ic["ind_ineq"], ic["ind_lb"], ic["ind_ub"]
        self.ind_llb, self.ind_uub = ic["ind_llb"], ic["ind_uub"]
        self.nx, self.ns = qp.nvar, len(self.ind_ineq)
        self.n, self.m = self.nx + self.ns, qp.ncon
        self.nlb, self.nub = len(self.ind_lb), len(self.ind_ub)
        n, m = self.n, self.m
        if self.opt.kkt_system == "Normal" and qp.nnzh > 0:
            # src/KKT/normalkkt.jl:45-48
            raise ValueError("NormalKKTSystem supports only linear programs")
        self.x, self.xl, self.xu = np.zeros(n), np.zeros(n), np.zeros(n)
        self.zl, self.zu, self.f = np.zeros(n), np.zeros(n), np.zeros(n)
        self.y, self.c, self.rhs, self.jacl = np.zeros(m), np.zeros(m), np.zeros(m), np.zeros(n)
        N = n + m + self.nlb + self.nub
        self.d, self.p, self._w1, self._w2 = np.zeros(N), np.zeros(N), np.zeros(N), np.zeros(N)
        self.correction_lb, self.correction_ub = np.zeros(self.nlb), np.zeros(self.nub)
        # kkt fields (normalkkt.jl:58-67 / SparseKKTSystem)
        self.reg, self.pr_diag, self.du_diag = np.zeros(n), np.zeros(n), np.zeros(m)
        self.l_diag, self.u_diag = np.zeros(self.nlb), np.zeros(self.nub)
        self.l_lower, self.u_lower = np.zeros(self.nlb), np.zeros(self.nub)
        self.obj_val = 0.0
        self.inf_pr = self.inf_du = self.inf_compl = 0.0
        self.norm_b = self.norm_c = 0.0
        self.mu = self.mu_curr = 0.0
        self.alpha_p = self.alpha_d = 0.0
        self.del_w = self.del_c = 0.0
        self.best_complementarity = np.inf
        self.status = "INITIAL"
        self.k = 0
        self.trace = []
        self.linear_solver_time = 0.0
        self.timers = dict(assembly=0.0, factor=0.0, solve=0.0, other=0.0)
        self.obj_scale = 1.0
        self.con_scale = np.ones(m)
        self._reg_state = list(self.opt.regularization)
        if self.opt.linear_solver == "auto":
            self.opt.linear_solver = "ldl" if self.opt.kkt_system == "Normal" else "splu"
        self.scaling_factor = np.ones(n)
        self._build_model()
        self._create_kkt_system()

    # ---------------------------------------------------------------- model callbacks
    def _build_model(self):
        """QuadraticModel callbacks through MadNLP.SparseCallback: A gets slack columns
        (-1 on inequality rows), H is expanded to a full symmetric operator."""
        qp, n, m, nx = self.qp, self.n, self.m, self.nx
        I = np.concatenate([qp.Arows, self.ind_ineq]).astype(np.int64)
        J = np.concatenate([qp.Acols, nx + np.arange(self.ns)]).astype(np.int64)
        V = np.concatenate([qp.Avals, -np.ones(self.ns)])
        self.A_I, self.A_J, self.A_V = I, J, V
        self.A = sp.csr_matrix((V, (I, J)), shape=(m, n))
        self.As = self.A                # scaled Jacobian (set_scaling!); identical while con_scale == 1
        Hl = sp.csr_matrix((qp.Hvals, (qp.Hrows, qp.Hcols)), shape=(nx, nx))
        self.Hfull = (Hl + sp.tril(Hl, -1).T).tocsr()
        self.cvec = np.concatenate([qp.c, np.zeros(self.ns)])

    def _eval_f(self, x):
        xv = x[: self.nx]
        return self.obj_scale * (self.qp.c0 + self.qp.c @ xv + 0.5 * (xv @ (self.Hfull @ xv)))

    def _eval_grad(self, x):
        g = np.zeros(self.n)
        xv = x[: self.nx]
        g[: self.nx] = self.obj_scale * (self.Hfull @ xv + self.qp.c)
        return g

    def _scaled_jacobian(self):
        """Jacobian with MadNLP's constraint scaling: the model entries of row i times con_scale[i]; the slack entries
        stay -1 (normalkkt.jl:163-172 / MadNLP.set_scaling!: the slack VARIABLES are scaled instead)."""
        cs = np.concatenate([self.con_scale[self.A_I[:len(self.qp.Avals)]], np.ones(self.ns)])
        return self.A_V * cs

    def _eval_cons(self, x):
        """c(x) = A x - s - rhs with the slack columns inside self.A (App. A)."""
        return self.As @ x - self.rhs

    def _jtprod(self, y):
        return self.As.T @ y

    def _jprod(self, x):
        return self.As @ x

    # ---------------------------------------------------------------- KKT systems
    def _create_kkt_system(self):
        n, m = self.n, self.m
        if self.opt.kkt_system == "Normal":
            # src/KKT/normalkkt.jl:70-115
            Ap, Aj, Ax = sparse_ref.coo_to_csr(m, n, self.A_I, self.A_J, np.arange(len(self.A_I), dtype=np.float64))
            self.A_csr_map = Ax.astype(np.int64)
            self.AT_p, self.AT_j = Ap, Aj
            self.AT_x = np.zeros(len(Aj))
            if self.opt.fast_symbolic:
                pat = sp.csr_matrix((np.ones(len(Aj)), Aj, Ap), shape=(m, n))
                low = sp.tril(pat @ pat.T).tocsc()
                low.sort_indices()
                self.C_p, self.C_j = low.indptr.astype(np.int32), low.indices.astype(np.int32)
            else:
                self.C_p, self.C_j = sparse_ref.build_normal_system(m, n, Ap, Aj)
            self.C_x = np.zeros(len(self.C_j))
            self.ls = _LinearSolver(m, self.C_p, self.C_j, self.opt.linear_solver, self.opt.ordering)
        elif self.opt.kkt_system in ("K2", "K2.5"):
            # MadNLP.SparseKKTSystem / ScaledSparseKKTSystem (App. B): same pattern, K2.5 = S K2 S: COO [pr_diag; hess; jac; slack; du_diag], lower triangular
            qp = self.qp
            I = np.concatenate([np.arange(n), qp.Hrows, n + self.A_I, n + np.arange(m)]).astype(np.int64)
            J = np.concatenate([np.arange(n), qp.Hcols, self.A_J, n + np.arange(m)]).astype(np.int64)
            self.aug_I, self.aug_J = I, J
            dim = n + m
            # SparseArrays.sparse pattern: column-major, rows sorted, duplicates merged
            key = J * dim + I
            uniq, inv = np.unique(key, return_inverse=True)
            self.aug_map = inv.astype(np.int64)
            rowval = (uniq % dim).astype(np.int32)
            colidx = uniq // dim
            colptr = np.zeros(dim + 1, dtype=np.int32)
            np.add.at(colptr, colidx + 1, 1)
            colptr = np.cumsum(colptr).astype(np.int32)
            self.aug_colptr, self.aug_rowval = colptr, rowval
            self.aug_nz = np.zeros(len(rowval))
            self.ls = _LinearSolver(dim, colptr, rowval, self.opt.linear_solver, self.opt.ordering)
        else:
            raise ValueError(self.opt.kkt_system)

    def _compress_jacobian(self):
        """normalkkt.jl:163-172 (Normal); values scaled by con_scale like MadNLP's jac callback."""
        V = self._scaled_jacobian()
        self.jac_V = V
        if self.opt.kkt_system == "Normal":
            self.AT_x = V[self.A_csr_map]

    def _build_kkt(self):
        t0 = time.perf_counter()
        if self.opt.kkt_system == "Normal":
            # normalkkt.jl:180-194
            D = 1.0 / self.pr_diag
            self.C_x = sparse_ref.assemble_normal_system(self.m, self.n, self.AT_p, self.AT_j, self.AT_x,
                                                         self.C_p, self.C_j, D)
        else:
            hess = self.obj_scale * self.qp.Hvals
            jac = self.jac_V
            if self.opt.kkt_system == "K2.5":       # ScaledSparseKKTSystem: S W S, J S
                sf = self.scaling_factor
                hess = hess * sf[self.qp.Hrows] * sf[self.qp.Hcols]
                jac = jac * sf[self.A_J]
            V = np.concatenate([self.pr_diag, hess, jac, self.du_diag])
            self.aug_nz = sparse_ref.transfer(len(self.aug_rowval), V, self.aug_map)
        self.timers["assembly"] += time.perf_counter() - t0

    def _factorize_wrapper(self):
        """MadNLP.factorize_wrapper!: build_kkt! then factorize! (timed)."""
        self._build_kkt()
        t0 = time.perf_counter()
        nz = self.C_x if self.opt.kkt_system == "Normal" else self.aug_nz
        self.ls.factorize(nz)
        dt = time.perf_counter() - t0
        self.linear_solver_time += dt
        self.timers["factor"] += dt

    # views of the unreduced KKT vector [xp(n); y(m); zl(nlb); zu(nub)]  (structure.jl:132)
    def _split(self, v):
        n, m, nlb = self.n, self.m, self.nlb
        return v[:n], v[n:n + m], v[n + m:n + m + nlb], v[n + m + nlb:]

    def _kkt_solve(self, w):
        """solve!(kkt, w): reduce_rhs! -> reduced solve -> finish_aug_solve!.
        Normal: normalkkt.jl:196-219; K2: MadNLP (App. B)."""
        wx, wy, wzl, wzu = self._split(w)
        if self.opt.kkt_system == "K2.5":
            # solve!(::ScaledSparseKKTSystem): l_diag = x - xl, u_diag = xu - x are positive; the primal block is scaled by S
            sf = self.scaling_factor
            np.add.at(wx, self.ind_lb, wzl / self.l_diag)
            np.add.at(wx, self.ind_ub, wzu / self.u_diag)
            t0 = time.perf_counter()
            sol = self.ls.solve(np.concatenate([wx * sf, wy]))
            self.timers["solve"] += time.perf_counter() - t0
            wx[:] = sol[: self.n] * sf
            wy[:] = sol[self.n:]
            wzl[:] = (wzl - self.l_lower * wx[self.ind_lb]) / self.l_diag
            wzu[:] = (-wzu + self.u_lower * wx[self.ind_ub]) / self.u_diag
            return w
        # reduce_rhs!
        np.subtract.at(wx, self.ind_lb, wzl / self.l_diag)
        np.subtract.at(wx, self.ind_ub, wzu / self.u_diag)
        t0 = time.perf_counter()
        if self.opt.kkt_system == "Normal":
            Sigma = self.pr_diag
            r1 = wx / Sigma
            r2 = self._jprod(r1) - wy
            dy = self.ls.solve(r2)
            wy[:] = dy
            r1 = wx - self._jtprod(wy)
            wx[:] = r1 / Sigma
        else:
            sol = self.ls.solve(np.concatenate([wx, wy]))
            wx[:] = sol[: self.n]
            wy[:] = sol[self.n:]
        self.timers["solve"] += time.perf_counter() - t0
        # finish_aug_solve!
        wzl[:] = (-wzl + self.l_lower * wx[self.ind_lb]) / self.l_diag
        wzu[:] = (wzu - self.u_lower * wx[self.ind_ub]) / self.u_diag
        return w

    def _kkt_mul(self, w, v, alpha, beta):
        """mul!(w, kkt, v, alpha, beta): normalkkt.jl:221-233 + MadNLP._kktmul! (App. B).
        For K2 the (1,1) block also carries the Hessian."""
        wx, wy, wzl, wzu = self._split(w)
        vx, vy, vzl, vzu = self._split(v)
        Hx = 0.0
        if self.opt.kkt_system in ("K2", "K2.5") and self.qp.nnzh > 0:
            Hx = np.zeros(self.n)
            Hx[: self.nx] = self.obj_scale * (self.Hfull @ vx[: self.nx])
        wx[:] = alpha * (self._jtprod(vy) + Hx) + beta * wx
        wy[:] = alpha * self._jprod(vx) + beta * wy
        wx += alpha * self.reg * vx
        wy += alpha * self.du_diag * vy
        np.subtract.at(wx, self.ind_lb, alpha * vzl)
        np.add.at(wx, self.ind_ub, alpha * vzu)
        sgn = -1.0 if self.opt.kkt_system == "K2.5" else 1.0       # K2.5 keeps l_diag, u_diag with the opposite sign
        wzl[:] = beta * wzl + alpha * (vx[self.ind_lb] * self.l_lower - sgn * vzl * self.l_diag)
        wzu[:] = beta * wzu + alpha * (vx[self.ind_ub] * self.u_lower + sgn * vzu * self.u_diag)
        return w

    def _solve_system(self):
        """src/linear_solver.jl:19-44."""
        self.d[:] = self.p
        self._kkt_solve(self.d)
        w = self._w1
        w[:] = self.p
        self._kkt_mul(w, self.d, -1.0, 1.0)
        norm_w = np.linalg.norm(w, np.inf)
        norm_p = np.linalg.norm(self.p, np.inf)
        self.residual_ratio = norm_w / max(1.0, norm_p)
        if np.isnan(self.residual_ratio) or (self.opt.check_residual and self.residual_ratio > self.opt.tol_linear_solve):
            raise FloatingPointError("SolveException")
        return self.d

    # ---------------------------------------------------------------- regularization (kernels.jl:364-401)
    def _init_regularization(self):
        kind = self._reg_state[0]
        self.del_w = 1.0
        self.del_c = 0.0 if kind == "none" else self._reg_state[2]

    def _update_regularization(self):
        r = self._reg_state
        if r[0] == "none":
            self.del_w, self.del_c = 0.0, 0.0
        elif r[0] == "fixed":
            self.del_w, self.del_c = r[1], r[2]
        elif r[0] == "adaptive":
            r[1] = max(r[1] / 10.0, r[3])
            r[2] = min(r[2] / 10.0, -r[3])
            self.del_w, self.del_c = r[1], r[2]
        else:
            raise ValueError(r[0])

    def _set_aug_diagonal_reg(self):
        """src/kernels.jl:124-136; ScaledSparseKKTSystem (K2.5): src/kernels.jl:139-149 + MadNLP._set_aug_diagonal!."""
        self.reg[:] = self.del_w
        self.du_diag[:] = self.del_c
        if self.opt.kkt_system == "K2.5":
            self.l_diag[:] = self.x[self.ind_lb] - self.xl[self.ind_lb]      # (X - Xl), positive
            self.u_diag[:] = self.xu[self.ind_ub] - self.x[self.ind_ub]      # (Xu - X), positive
            self.l_lower[:] = self.zl[self.ind_lb]
            self.u_lower[:] = self.zu[self.ind_ub]
            xlzu, xuzl = np.zeros(self.n), np.zeros(self.n)
            xlzu[self.ind_ub] = self.u_lower
            xlzu[self.ind_lb] *= self.l_diag
            xuzl[self.ind_lb] = self.l_lower
            xuzl[self.ind_ub] *= self.u_diag
            self.pr_diag[:] = xlzu + xuzl
            sf = np.ones(self.n)
            sf[self.ind_lb] *= np.sqrt(self.l_diag)
            sf[self.ind_ub] *= np.sqrt(self.u_diag)
            self.scaling_factor = sf
            self.pr_diag += self.reg * sf ** 2
            return
        self.l_diag[:] = self.xl[self.ind_lb] - self.x[self.ind_lb]
        self.u_diag[:] = self.x[self.ind_ub] - self.xu[self.ind_ub]
        self.l_lower[:] = self.zl[self.ind_lb]
        self.u_lower[:] = self.zu[self.ind_ub]
        self.pr_diag[:] = self.reg
        self.pr_diag[self.ind_lb] -= self.l_lower / self.l_diag
        self.pr_diag[self.ind_ub] -= self.u_lower / self.u_diag

    def _factorize_regularized_system(self):
        """src/linear_solver.jl:6-17."""
        for _ in range(3):
            self._set_aug_diagonal_reg()
            self._factorize_wrapper()
            if self.ls.ok:
                break
            self.del_w *= 100.0
            self.del_c *= 100.0

    # ---------------------------------------------------------------- initialization
    def _madnlp_initialize(self):
        """MadNLP.initialize!(cb, x, xl, xu, y0, rhs, ind_ineq; tol, bound_push, bound_fac) (App. B)."""
        qp, opt, nx = self.qp, self.opt, self.nx
        self.x[:nx] = qp.x0
        self.x[nx:] = 0.0
        self.y[:] = qp.y0
        self.xl[:nx], self.xu[:nx] = qp.lvar, qp.uvar
        self.xl[nx:], self.xu[nx:] = qp.lcon[self.ind_ineq], qp.ucon[self.ind_ineq]
        self.rhs[:] = np.where(qp.lcon == qp.ucon, qp.lcon, 0.0)
        tol = opt.bound_relax_factor
        self.xl -= np.maximum(1.0, np.abs(self.xl)) * tol     # -inf stays -inf
        self.xu += np.maximum(1.0, np.abs(self.xu)) * tol
        # _initialize_variables!
        x, xl, xu = self.x, self.xl, self.xu
        bp, bf = opt.bound_push, opt.bound_fac
        fl, fu = np.isfinite(xl), np.isfinite(xu)
        both, lo, up = fl & fu, fl & ~fu, ~fl & fu
        with np.errstate(invalid="ignore"):
            pl = np.minimum(bp * np.maximum(1.0, np.abs(xl)), bf * (xu - xl))
            pu = np.minimum(bp * np.maximum(1.0, np.abs(xu)), bf * (xu - xl))
            x[both] = np.maximum(xl + pl, np.minimum(xu - pu, x))[both]
            x[lo] = np.maximum(xl + bp * np.maximum(1.0, np.abs(xl)), x)[lo]
            x[up] = np.minimum(xu - bp * np.maximum(1.0, np.abs(xu)), x)[up]

    def _set_scaling(self):
        """MadNLP.set_scaling!(..., 100) (App. B): identity when |A_ij|, ||grad f(x0)|| <= 100."""
        maxg = 100.0
        rowmax = np.zeros(self.m)
        np.maximum.at(rowmax, self.A_I, np.abs(self.A_V))
        self.con_scale = np.minimum(1.0, maxg / np.maximum(rowmax, 1e-300))
        # MadNLP.set_scaling! (MadNLP 0.8, un-vendored): y0 ./= con_scale; rhs .*= con_scale; slack(x), slack(xl),
        # slack(xu) .*= con_scale[ind_ineq] (the slack columns of the Jacobian stay -1)
        self.y /= self.con_scale
        self.rhs *= self.con_scale
        cs = self.con_scale[self.ind_ineq]
        self.x[self.nx:] *= cs
        self.xl[self.nx:] *= cs
        self.xu[self.nx:] *= cs
        self.As = sp.csr_matrix((self._scaled_jacobian(), (self.A_I, self.A_J)), shape=(self.m, self.n))
        self.obj_scale = 1.0
        g = np.linalg.norm(self._eval_grad(self.x), np.inf)
        self.obj_scale = min(1.0, maxg / g) if g > 0 else 1.0

    def _kkt_initialize(self):
        """MadNLP.initialize!(kkt): normalkkt.jl:150-161."""
        self.reg[:] = 1.0
        self.pr_diag[:] = 1.0
        self.du_diag[:] = 0.0
        self.l_lower[:] = 0.0
        self.u_lower[:] = 0.0
        self.l_diag[:] = 1.0
        self.u_diag[:] = 1.0
        self.scaling_factor = np.ones(self.n)

    def initialize(self):
        """src/solver.jl:127-189."""
        opt = self.opt
        self._madnlp_initialize()
        self.jacl[:] = 0.0
        if opt.scaling:
            self._set_scaling()
        self._kkt_initialize()
        self._init_regularization()
        self.obj_val = self._eval_f(self.x)
        self._compress_jacobian()
        self.f[:] = self._eval_grad(self.x)
        self.c[:] = self._eval_cons(self.x)
        self.norm_b = np.linalg.norm(self.rhs, np.inf) if self.m else 0.0
        self.norm_c = np.linalg.norm(self.f, np.inf)
        self.init_starting_point()
        self.mu = opt.mu_init
        self.best_complementarity = np.inf
        self.status = REGULAR
        self.jacl[:] = self._jtprod(self.y)

    def init_starting_point(self):
        """src/solver.jl:6-125."""
        lbi, ubi = self.ind_lb, self.ind_ub
        x, l, u = self.x, self.xl, self.xu
        n, m = self.n, self.m
        self.reg[:] = self.del_w
        self.pr_diag[:] = self.del_w
        self.du_diag[:] = self.del_c
        self._factorize_wrapper()
        # Step 1 (kernels.jl:1-9)
        self.p[:] = 0.0
        self.p[n:n + m] = -self.c
        self._solve_system()
        x += self.d[:n]
        # Step 2 (kernels.jl:11-19)
        self.p[:] = 0.0
        self.p[:n] = -self.f
        self._solve_system()
        self.y[:] = self.d[n:n + m]
        # Step 3
        res = self._jtprod(self.y) + self.f
        fl, fu = np.isfinite(l), np.isfinite(u)
        self.zl[:] = np.where(fl & fu, 0.5 * res, np.where(fl, res, self.zl))
        self.zu[:] = np.where(fl & fu, -0.5 * res, np.where(fu, -res, self.zu))

        def _min0(a):
            return min(0.0, a.min()) if len(a) else 0.0

        delta_x = max(0.0, -1.5 * _min0(x[lbi] - l[lbi]), -1.5 * _min0(u[ubi] - x[ubi]))
        delta_s = max(0.0, -1.5 * _min0(self.zl[lbi]), -1.5 * _min0(self.zu[ubi]))
        x[lbi] = x[lbi] + delta_x
        x[ubi] = x[ubi] - delta_x
        self.zl[lbi] += 1.0 + delta_s
        self.zu[ubi] += 1.0 + delta_s
        mu = 0.0
        if self.nlb > 0:
            mu += x[lbi] @ self.zl[lbi] - l[lbi] @ self.zl[lbi]
        if self.nub > 0:
            mu += u[ubi] @ self.zu[ubi] - x[ubi] @ self.zu[ubi]
        with np.errstate(divide="ignore", invalid="ignore"):
            delta_x2 = np.float64(mu) / (2 * (self.zl[lbi].sum() + self.zu[ubi].sum()))
            delta_s2 = np.float64(mu) / (2 * ((x[lbi] - l[lbi]).sum() + (u[ubi] - x[ubi]).sum()))
        x[lbi] += delta_x2
        x[ubi] -= delta_x2
        self.zl[lbi] += delta_s2
        self.zu[ubi] += delta_s2
        kappa = self.opt.bound_fac
        with np.errstate(invalid="ignore"):
            below, above = x < l, u < x
            pl = np.minimum(kappa * np.maximum(1.0, l), kappa * (u - l))
            pu = np.minimum(kappa * np.maximum(1.0, u), kappa * (u - l))
            x[:] = np.where(below, l + pl, np.where(above, u - pu, x))
        assert np.all(self.zl[lbi] > 0.0) and np.all(self.zu[ubi] > 0.0)
        assert np.all(x[lbi] > l[lbi]) and np.all(x[ubi] < u[ubi])

    # ---------------------------------------------------------------- MPC pieces
    def _dual_objective(self):
        """src/kernels.jl:408-417."""
        dobj = -(self.y @ self.rhs)
        if self.nlb > 0:
            dobj += self.zl[self.ind_lb] @ self.xl[self.ind_lb]
        if self.nub > 0:
            dobj -= self.zu[self.ind_ub] @ self.xu[self.ind_ub]
        return dobj

    def _get_optimality_gap(self):
        """src/kernels.jl:419-430 + MadNLP.get_inf_compl with mu=0, sc=1."""
        lbi, ubi = self.ind_lb, self.ind_ub
        a = np.abs((self.x[lbi] - self.xl[lbi]) * self.zl[lbi]).max() if self.nlb else 0.0
        b = np.abs((self.xu[ubi] - self.x[ubi]) * self.zu[ubi]).max() if self.nub else 0.0
        return max(a, b)

    def update_termination_criteria(self):
        """src/solver.jl:194-222."""
        opt = self.opt
        self.dobj = self._dual_objective()
        self.inf_pr = (np.linalg.norm(self.c, np.inf) if self.m else 0.0) / max(1.0, self.norm_b)
        self.inf_du = np.linalg.norm(self.f - self.zl + self.zu + self.jacl, np.inf) / max(1.0, self.norm_c)
        self.inf_compl = self._get_optimality_gap() / max(1.0, self.norm_c)
        self.best_complementarity = min(self.best_complementarity, self.inf_compl)
        if max(self.inf_pr, self.inf_du, self.inf_compl) <= opt.tol:
            self.status = SOLVE_SUCCEEDED
        elif (self.inf_compl > opt.divergence_tol * self.best_complementarity) and \
                (self.dobj > max(10.0 * abs(self.obj_val), 1.0)):
            self.status = INFEASIBLE_PROBLEM_DETECTED
        elif self.obj_val < -opt.divergence_tol * max(10.0, abs(self.dobj), 1.0):
            self.status = DIVERGING_ITERATES
        elif self.k >= opt.max_iter:
            self.status = MAXIMUM_ITERATIONS_EXCEEDED
        elif time.time() - self.start_time >= opt.max_wall_time:
            self.status = MAXIMUM_WALLTIME_EXCEEDED

    def _set_predictive_rhs(self):
        """src/kernels.jl:21-41."""
        n, m, nlb = self.n, self.m, self.nlb
        lbi, ubi = self.ind_lb, self.ind_ub
        p = self.p
        p[:] = 0.0
        p[:n] = -self.f + self.zl - self.zu - self.jacl
        p[n:n + m] = -self.c
        p[n + m:n + m + nlb] = (self.xl[lbi] - self.x[lbi]) * self.zl[lbi]
        p[n + m + nlb:] = (self.xu[ubi] - self.x[ubi]) * self.zu[ubi]

    def _set_correction_rhs(self, mu):
        """src/kernels.jl:43-58."""
        n, m, nlb = self.n, self.m, self.nlb
        lbi, ubi = self.ind_lb, self.ind_ub
        p = self.p
        p[:n] = -self.f + self.zl - self.zu - self.jacl
        p[n:n + m] = -self.c
        p[n + m:n + m + nlb] = (self.xl[lbi] - self.x[lbi]) * self.zl[lbi] + mu - self.correction_lb
        p[n + m + nlb:] = (self.xu[ubi] - self.x[ubi]) * self.zu[ubi] - mu - self.correction_ub

    def _get_correction(self):
        """src/kernels.jl:60-71."""
        dx, _, dzl, dzu = self._split(self.d)
        self.correction_lb[:] = dx[self.ind_lb] * dzl
        self.correction_ub[:] = dx[self.ind_ub] * dzu

    def _set_extra_correction(self, alpha_p, alpha_d, bmin, bmax, mu):
        """src/kernels.jl:74-122."""
        dx, _, dzl, dzu = self._split(self.d)
        lbi, ubi = self.ind_lb, self.ind_ub
        tmin, tmax = bmin * mu, bmax * mu
        v = (self.x[lbi] + alpha_p * dx[lbi] - self.xl[lbi]) * (self.zl[lbi] + alpha_d * dzl)
        delta = np.where(v < tmin, tmin - v, np.where(v > tmax, tmax - v, 0.0))
        self.correction_lb -= delta
        v = (self.xu[ubi] - alpha_p * dx[ubi] - self.x[ubi]) * (self.zu[ubi] + alpha_d * dzu)
        delta = np.where(v < tmin, tmin - v, np.where(v > tmax, tmax - v, 0.0))
        self.correction_ub += delta

    def _get_complementarity_measure(self):
        """src/kernels.jl:155-174."""
        if self.nlb + self.nub == 0:
            return 0.0
        lbi, ubi = self.ind_lb, self.ind_ub
        a = ((self.x[lbi] - self.xl[lbi]) * self.zl[lbi]).sum()
        b = ((self.xu[ubi] - self.x[ubi]) * self.zu[ubi]).sum()
        return (a + b) / (self.nlb + self.nub)

    def _get_affine_complementarity_measure(self, alpha_p, alpha_d):
        """src/kernels.jl:176-208."""
        if self.nlb + self.nub == 0:
            return 0.0
        dx, _, dzl, dzu = self._split(self.d)
        lbi, ubi = self.ind_lb, self.ind_ub
        a = (((self.x[lbi] + alpha_p * dx[lbi]) - self.xl[lbi]) * (self.zl[lbi] + alpha_d * dzl)).sum()
        b = ((self.xu[ubi] - (self.x[ubi] + alpha_p * dx[ubi])) * (self.zu[ubi] + alpha_d * dzu)).sum()
        return (a + b) / (self.nlb + self.nub)

    def _update_barrier(self, mu_affine):
        """src/kernels.jl:210-220 (Mehrotra). Quirk A.9(vii): has_inequalities uses the swapped
        fields, i.e. (nlb + nub) > 0."""
        has_ineq = (self.nlb + self.nub) > 0
        mu_curr = self._get_complementarity_measure()
        sigma = float(np.clip((mu_affine / mu_curr) ** 3, 1e-6, 10.0)) if has_ineq else 1.0
        self.mu = max(self.opt.mu_min, sigma * mu_curr)
        return mu_curr

    @staticmethod
    def _argmin_ratio(val):
        """mapreduce with init (1.0, 0) and op (a, b) -> a[1] < b[1] ? a : b (kernels.jl:227-269):
        a left fold, so on ties the RIGHT-most element wins, and an element equal to the init
        value 1.0 beats the init."""
        best, idx = 1.0, 0
        if len(val):
            j = len(val) - 1 - int(np.argmin(val[::-1]))
            if not (best < val[j]):
                best, idx = float(val[j]), j + 1
        return best, idx

    def _get_alpha_max_primal(self, tau):
        """src/kernels.jl:226-248."""
        dx = self.d[: self.n]
        lbi, ubi = self.ind_lb, self.ind_ub
        dxl, dxu = dx[lbi], dx[ubi]
        with np.errstate(divide="ignore", invalid="ignore"):
            vl = np.where(dxl < 0, (-self.x[lbi] + self.xl[lbi]) * tau / dxl, np.inf)
            vu = np.where(dxu > 0, (-self.x[ubi] + self.xu[ubi]) * tau / dxu, np.inf)
        axl, il = self._argmin_ratio(vl)
        axu, iu = self._argmin_ratio(vu)
        return axl, axu, il, iu

    def _get_alpha_max_dual(self, tau):
        """src/kernels.jl:250-272 (note the asymmetric zu condition, quirk A.9 iii)."""
        _, _, dzl, dzu = self._split(self.d)
        zl, zu = self.zl[self.ind_lb], self.zu[self.ind_ub]
        with np.errstate(divide="ignore", invalid="ignore"):
            vl = np.where(dzl < 0, (-zl) * tau / dzl, np.inf)
            vu = np.where((dzu < 0) & (zu + dzu < 0), (-zu) * tau / dzu, np.inf)
        azl, il = self._argmin_ratio(vl)
        azu, iu = self._argmin_ratio(vu)
        return azl, azu, il, iu

    def _get_fraction_to_boundary_step(self, tau):
        """src/kernels.jl:274-289."""
        axl, axu, _, _ = self._get_alpha_max_primal(tau)
        azl, azu, _, _ = self._get_alpha_max_dual(tau)
        return min(axl, axu), min(azl, azu)

    def _update_step(self):
        """src/kernels.jl:291-358."""
        rule = self.opt.step_rule
        if rule[0] == "conservative":
            self.alpha_p, self.alpha_d = self._get_fraction_to_boundary_step(rule[1])
        elif rule[0] == "adaptive":
            tau = max(1 - self.mu, rule[1])
            self.alpha_p, self.alpha_d = self._get_fraction_to_boundary_step(tau)
        elif rule[0] == "mehrotra":
            gamma_f = rule[1]
            gamma_a = 1.0 / (1.0 - gamma_f)
            dx, _, dzl, dzu = self._split(self.d)
            lbi, ubi = self.ind_lb, self.ind_ub
            axl, axu, i_xl, i_xu = self._get_alpha_max_primal(1.0)
            azl, azu, i_zl, i_zu = self._get_alpha_max_dual(1.0)
            max_ap, max_ad = min(axl, axu), min(azl, azu)
            mu_full = self._get_affine_complementarity_measure(max_ap, max_ad) / gamma_a
            ap, ad = 1.0, 1.0
            if max_ap < 1.0:
                if axl <= axu:
                    i = i_xl - 1
                    tmp = mu_full / (self.zl[lbi][i] + max_ad * dzl[i])
                    ap = (self.x[lbi][i] - self.xl[lbi][i] - tmp) / (-dx[lbi][i])
                else:
                    i = i_xu - 1
                    tmp = mu_full / (self.zu[ubi][i] + max_ad * dzu[i])
                    ap = (self.xu[ubi][i] - self.x[ubi][i] - tmp) / (dx[ubi][i])
            if max_ad < 1.0:
                if azl <= azu:
                    i = i_zl - 1
                    tmp = mu_full / (self.x[lbi][i] + max_ap * dx[lbi][i] - self.xl[lbi][i])
                    ad = -(self.zl[lbi][i] - tmp) / dzl[i]
                else:
                    i = i_zu - 1
                    tmp = mu_full / (self.xu[ubi][i] - self.x[ubi][i] - max_ap * dx[ubi][i])
                    ad = -(self.zu[ubi][i] - tmp) / dzu[i]
            self.alpha_p = max(ap, gamma_f * max_ap)
            self.alpha_d = max(ad, gamma_f * max_ad)
        else:
            raise ValueError(rule[0])

    def _gondzio(self):
        """src/solver.jl:245-298."""
        if self.opt.max_ncorr <= 0:
            return
        delta, bmin, bmax, tau = 0.1, 0.1, 10.0, 0.995
        alpha_p, alpha_d = self._get_fraction_to_boundary_step(tau)
        for _ in range(self.opt.max_ncorr):
            tap, tad = min(alpha_p + delta, 1.0), min(alpha_d + delta, 1.0)
            ga = self._get_affine_complementarity_measure(tap, tad)
            g = self.mu_curr
            mu = (ga / g) ** 2 * ga
            self._set_extra_correction(tap, tad, bmin, bmax, mu)
            self._set_correction_rhs(mu)
            self._w2[:] = self.d
            self._solve_system()
            hap, had = self._get_fraction_to_boundary_step(tau)
            if (hap < 1.005 * alpha_p) or (had < 1.005 * alpha_d):
                self.d[:] = self._w2
                break
            alpha_p, alpha_d = hap, had

    def _adjust_boundary(self):
        """MadNLP.adjust_boundary! (App. B)."""
        eps = np.finfo(np.float64).eps
        c1, c2 = eps * self.mu, eps ** 0.75
        lbi, ubi = self.ind_lb, self.ind_ub
        xl_r, xu_r = self.xl[lbi], self.xu[ubi]
        x_lr, x_ur = self.x[lbi], self.x[ubi]
        self.xl[lbi] = np.where(x_lr - xl_r < c1, xl_r - c2 * np.maximum(1.0, np.abs(x_lr)), xl_r)
        self.xu[ubi] = np.where(xu_r - x_ur < c1, xu_r + c2 * np.maximum(1.0, np.abs(x_ur)), xu_r)

    def _apply_step(self):
        """src/solver.jl:308-317."""
        dx, dy, dzl, dzu = self._split(self.d)
        self.x += self.alpha_p * dx
        self.y += self.alpha_d * dy
        self.zl[self.ind_lb] += self.alpha_d * dzl
        self.zu[self.ind_ub] += self.alpha_d * dzu
        self._adjust_boundary()
        self.k += 1

    def _evaluate_model(self):
        """src/solver.jl:319-326."""
        self.obj_val = self._eval_f(self.x)
        self.c[:] = self._eval_cons(self.x)
        self.f[:] = self._eval_grad(self.x)
        self.jacl[:] = self._jtprod(self.y)

    def _record(self):
        self.trace.append(dict(
            k=self.k, objective=self.obj_val / self.obj_scale, dual_objective=self.dobj / self.obj_scale,
            inf_pr=self.inf_pr, inf_du=self.inf_du, inf_compl=self.inf_compl, mu=self.mu,
            alpha_p=self.alpha_p, alpha_d=self.alpha_d, del_w=self.del_w,
            dnorm=0.0 if self.k == 0 else float(np.linalg.norm(self.d[: self.n], np.inf)),
        ))

    def mpc_iteration(self):
        """One pass of the loop body of mpc! (src/solver.jl:333-359). Returns False when done."""
        self.update_termination_criteria()
        self._record()
        if self.status != REGULAR:
            return False
        self._update_regularization()
        self._factorize_regularized_system()
        # prediction_step! (solver.jl:230-237)
        self._set_predictive_rhs()
        self._solve_system()
        ap_aff, ad_aff = self._get_fraction_to_boundary_step(1.0)
        mu_affine = self._get_affine_complementarity_measure(ap_aff, ad_aff)
        self._get_correction()
        self.mu_curr = self._update_barrier(mu_affine)
        # mehrotra_correction_direction! (solver.jl:239-243)
        self._set_correction_rhs(self.mu)
        self._solve_system()
        self._gondzio()
        self._update_step()
        self._apply_step()
        self._evaluate_model()
        return True

    def mpc(self):
        """src/solver.jl:332-360."""
        while self.mpc_iteration():
            pass

    def solve(self):
        """src/solver.jl:362-418."""
        self.start_time = time.time()
        t0 = time.perf_counter()
        try:
            self.initialize()
            self.mpc()
        except (FloatingPointError, AssertionError):
            self.status = INTERNAL_ERROR
        total = time.perf_counter() - t0
        dobj = getattr(self, "dobj", float("nan"))
        return Stats(
            status=self.status, iter=self.k, objective=self.obj_val / self.obj_scale,
            dual_objective=dobj / self.obj_scale, solution=self.x[: self.nx].copy(),
            constraints=np.asarray(sp.csr_matrix((self.qp.Avals, (self.qp.Arows, self.qp.Acols)), shape=(self.m, self.nx)) @ self.x[: self.nx]),
            multipliers=self.y * self.con_scale / self.obj_scale,
            multipliers_L=self.zl[: self.nx] / self.obj_scale, multipliers_U=self.zu[: self.nx] / self.obj_scale,
            trace=self.trace, total_time=total, linear_solver_time=self.linear_solver_time,
            timers=dict(self.timers),
        )


def madipm(qp, **kwargs):
    """src/solver.jl:425-428."""
    return MPCOracle(qp, **kwargs).solve()
