"""CPU tests of the oracle itself: pinned on the reference's own fixture (simple_lp), cross-checked
with HiGHS and algebraic identities, and guarded by committed traces."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp
from scipy.optimize import linprog

from madipm_jl_b200.problems import random_sparse_lp, random_sparse_qp, simple_lp
from oracle import sparse_ref
from oracle.mpc_oracle import MPCOracle, madipm

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "traces.json")))


def test_simple_lp_reference_pin():
    """test/runtests.jl:144-198: objective 1.0, SOLVE_SUCCEEDED; Normal == K2 at 1e-6."""
    ref = madipm(simple_lp(), regularization=("none",))
    assert ref.status == "SOLVE_SUCCEEDED"
    assert abs(ref.objective - 1.0) < 1e-8
    k2 = madipm(simple_lp())
    nm = madipm(simple_lp(), kkt_system="Normal", linear_solver="ldl")
    for s in (k2, nm):
        assert s.status == "SOLVE_SUCCEEDED"
        assert abs(s.objective - ref.objective) < 1e-6
        assert np.allclose(s.solution, ref.solution, atol=1e-6)
        assert np.allclose(s.multipliers, ref.multipliers, atol=1e-6)
        assert np.allclose(s.constraints, ref.constraints, atol=1e-6)
    assert np.allclose(k2.solution, [0.5, 0.5], atol=1e-6)


@pytest.mark.parametrize("rule", [("adaptive", 0.99), ("conservative", 0.99), ("mehrotra", 0.99)])
def test_step_rules(rule):
    """test/runtests.jl:85-97: every step rule reaches SOLVE_SUCCEEDED."""
    qp = random_sparse_lp(40, 160, 5, 7, structure="uniform", ub_fraction=0.5)
    assert madipm(qp, step_rule=rule).status == "SOLVE_SUCCEEDED"


@pytest.mark.parametrize("reg", [("fixed", 1e-8, -1e-9), ("adaptive", 1e-8, -1e-9, 1e-9)])
def test_regularizations(reg):
    """test/runtests.jl:122-140: regularized LDL solves agree with the unregularized one at 1e-6."""
    qp = random_sparse_qp(30, 80, 4, 5, structure="window", window=10)
    ref = madipm(qp, regularization=("none",))
    s = madipm(qp, regularization=reg)
    assert s.status == ref.status == "SOLVE_SUCCEEDED"
    assert abs(s.objective - ref.objective) < 1e-6 * max(1, abs(ref.objective))
    assert np.allclose(s.solution, ref.solution, atol=1e-5)


def test_gondzio():
    qp = random_sparse_lp(40, 160, 5, 7, structure="uniform", ub_fraction=0.5)
    a, b = madipm(qp), madipm(qp, max_ncorr=5)
    assert a.status == b.status == "SOLVE_SUCCEEDED"
    assert abs(a.objective - b.objective) < 1e-6 * max(1, abs(a.objective))


@pytest.mark.parametrize("case", [(40, 160, 5, "uniform", 0.5), (300, 1500, 5, "window", 0.0)])
def test_against_highs(case):
    m, n, k, structure, ubf = case
    qp = random_sparse_lp(m, n, k, 7, structure=structure, window=20, ub_fraction=ubf)
    A = sp.csr_matrix((qp.Avals, (qp.Arows, qp.Acols)), shape=(m, n))
    r = linprog(qp.c, A_eq=A, b_eq=qp.lcon, method="highs",
                bounds=list(zip(qp.lvar, [None if np.isinf(u) else u for u in qp.uvar])))
    for kkt in ("Normal", "K2"):
        s = madipm(qp, kkt_system=kkt)
        assert s.status == "SOLVE_SUCCEEDED"
        assert abs(s.objective - r.fun) <= 1e-6 * max(1.0, abs(r.fun))
        assert abs(s.objective - s.dual_objective) <= 1e-5 * max(1.0, abs(r.fun))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_mixed_bounds_against_highs(seed):
    """Range / one-sided rows (slack variables) and free / boxed / upper-only variables."""
    from madipm_jl_b200.problems import mixed_bounds_lp
    qp = mixed_bounds_lp(30, 90, 4, seed)
    m, n = qp.ncon, qp.nvar
    A = sp.csr_matrix((qp.Avals, (qp.Arows, qp.Acols)), shape=(m, n))
    eq = qp.lcon == qp.ucon
    fu, fl = np.isfinite(qp.ucon) & ~eq, np.isfinite(qp.lcon) & ~eq
    r = linprog(qp.c, A_ub=sp.vstack([A[fu], -A[fl]]), b_ub=np.concatenate([qp.ucon[fu], -qp.lcon[fl]]),
                A_eq=A[eq], b_eq=qp.lcon[eq], method="highs",
                bounds=[(None if np.isinf(l) else l, None if np.isinf(u) else u) for l, u in zip(qp.lvar, qp.uvar)])
    assert r.status == 0
    for kkt in ("K2", "Normal"):
        s = madipm(qp, kkt_system=kkt)
        assert s.status == "SOLVE_SUCCEEDED"
        assert abs(s.objective - r.fun) <= 1e-6 * max(1.0, abs(r.fun))
        assert np.all(s.constraints >= qp.lcon - 1e-6) and np.all(s.constraints <= qp.ucon + 1e-6)


def test_unbounded_lp_is_reported():
    """An LP with an unbounded ray must end as DIVERGING_ITERATES (src/solver.jl:212-213), not loop."""
    qp = random_sparse_lp(10, 30, 3, 1, structure="uniform")
    qp.c[:] = -np.abs(qp.c)                       # push along the recession cone of {Ax=b, x>=0}? not guaranteed:
    qp.lvar[:] = -np.inf                          # ... free variables make it unbounded for sure
    s = madipm(qp, max_iter=200)
    assert s.status in ("DIVERGING_ITERATES", "INFEASIBLE_PROBLEM_DETECTED", "MAXIMUM_ITERATIONS_EXCEEDED", "INTERNAL_ERROR")
    assert s.status != "SOLVE_SUCCEEDED"


def test_max_iter_status():
    s = madipm(random_sparse_lp(40, 160, 5, 7, structure="uniform"), max_iter=3)
    assert s.status == "MAXIMUM_ITERATIONS_EXCEEDED" and s.iter == 3


@pytest.mark.parametrize("key", sorted(GOLD))
def test_golden_traces(key):
    """The oracle reproduces its committed per-iterate traces (regression guard)."""
    from tests.golden.make_golden import CASES
    name, kkt = key.split("/")
    st = madipm(CASES[name](), kkt_system=kkt)
    g = GOLD[key]
    assert st.status == g["status"] and st.iter == g["iter"]
    for a, b in zip(st.trace, g["trace"]):
        for f in ("objective", "dual_objective", "inf_pr", "inf_du", "inf_compl", "mu", "alpha_p", "alpha_d"):
            assert abs(a[f] - b[f]) <= 1e-9 * max(1.0, abs(b[f])), (key, a["k"], f, a[f], b[f])


def test_unreduced_system_identity():
    """K*d == p for the unreduced Newton system after every solve of an iteration."""
    qp = random_sparse_lp(40, 160, 5, 3, structure="uniform", ub_fraction=0.5)
    for kkt in ("Normal", "K2"):
        o = MPCOracle(qp, kkt_system=kkt, regularization=("fixed", 1e-10, 0.0))
        o.start_time = 0.0
        o.initialize()
        o._update_regularization()
        o._factorize_regularized_system()
        o._set_predictive_rhs()
        o._solve_system()
        assert o.residual_ratio < 1e-9, (kkt, o.residual_ratio)


def test_normal_pattern_matches_sparse_product():
    """build_normal_system (utils.jl:209-274) == structural pattern of tril(A A')."""
    rng = np.random.default_rng(0)
    m, n = 70, 200
    A = sp.random(m, n, 0.03, random_state=1, format="coo")
    A.data[:] = 1.0
    Bp, Bj, Bx = sparse_ref.coo_to_csr(m, n, A.row, A.col, np.arange(A.nnz, dtype=float))
    Cp, Cj = sparse_ref.build_normal_system(m, n, Bp, Bj)
    P = sp.tril((A @ A.T).tocsr()).tocsc()
    P.sort_indices()
    # reference emits column i with rows j >= i  == lower CSC
    assert (P.indptr == Cp).all() and (P.indices == Cj).all()
    D = rng.uniform(0.5, 2.0, n)
    vals = rng.standard_normal(A.nnz)
    Cx = sparse_ref.assemble_normal_system(m, n, Bp, Bj, vals[Bx.astype(int)], Cp, Cj, D)
    Av = sp.csr_matrix((vals, (A.row, A.col)), shape=(m, n))
    ref = sp.tril(Av @ sp.diags(D) @ Av.T).tocsc()
    ref.sort_indices()
    dense = np.zeros((m, m))
    for i in range(m):
        dense[Cj[Cp[i]:Cp[i + 1]], i] = Cx[Cp[i]:Cp[i + 1]]
    assert np.allclose(dense, ref.toarray(), atol=1e-12)


def test_ldl_with_permutation():
    rng = np.random.default_rng(0)
    n = 40
    B = sp.random(n, n, 0.15, random_state=1)
    M = (B @ B.T + 2 * sp.eye(n)).toarray()
    low = sp.csc_matrix(np.tril(M))
    low.sort_indices()
    for perm in (None, rng.permutation(n).astype(np.int32)):
        L = sparse_ref.LDL(n, low.indptr.astype(np.int32), low.indices.astype(np.int32), perm)
        assert L.factorize(low.data) and L.inertia() == (n, 0, 0)
        b = rng.standard_normal(n)
        assert np.abs(M @ L.solve(b) - b).max() < 1e-12


@pytest.mark.parametrize("make", [simple_lp, lambda: random_sparse_qp(60, 200, 4, 9, structure="window", window=10),
                                  lambda: random_sparse_lp(40, 160, 5, 7, structure="uniform", ub_fraction=0.5)])
def test_k25_scaled_kkt_system_matches_k2(make):
    """test/runtests.jl:107-120: MadNLP.ScaledSparseKKTSystem (K2.5) must reproduce the K2 results (status, iterations,
    objective, solution, constraints, multipliers at 1e-6); simple_lp is the reference's own fixture."""
    qp = make()
    ref = madipm(qp, kkt_system="K2")
    k25 = madipm(qp, kkt_system="K2.5")
    assert k25.status == ref.status == "SOLVE_SUCCEEDED"
    assert k25.iter == ref.iter
    assert abs(k25.objective - ref.objective) <= 1e-6
    for f in ("solution", "constraints", "multipliers"):
        assert np.abs(getattr(k25, f) - getattr(ref, f)).max() <= 1e-6


def test_standard_form_qp_matches_the_original_problem():
    """test/runtests.jl:159-164: the standard-form reformulation (src/utils.jl:373-505) solves to the same objective."""
    from madipm_jl_b200.preprocess import standard_form_qp
    from madipm_jl_b200.problems import mixed_bounds_lp
    for qp in (mixed_bounds_lp(30, 90, 4, 1), random_sparse_qp(60, 200, 4, 9, structure="window", window=10),
               random_sparse_lp(40, 160, 5, 7, structure="uniform", ub_fraction=0.5)):
        std = standard_form_qp(qp)
        n_rng = int(np.sum(np.isfinite(qp.lvar) & np.isfinite(qp.uvar) & (qp.lvar < qp.uvar)))
        ineq = qp.lcon < qp.ucon
        n_rng += int(np.sum(np.isfinite(qp.lcon[ineq]) & np.isfinite(qp.ucon[ineq])))
        assert std.nvar == qp.nvar + int(ineq.sum()) + n_rng and std.ncon == qp.ncon + n_rng
        assert np.all(std.lcon == std.ucon)                   # only equality rows are left
        assert not np.any(np.isfinite(std.lvar) & np.isfinite(std.uvar) & (std.lvar < std.uvar) & (np.arange(std.nvar) < qp.nvar))
        a, b = madipm(qp, kkt_system="K2"), madipm(std, kkt_system="K2")
        assert a.status == b.status == "SOLVE_SUCCEEDED"
        assert abs(a.objective - b.objective) <= 1e-7 * max(1.0, abs(a.objective))


def test_ruiz_restatement_equilibrates():
    from oracle.preprocess_ref import ruiz_equilibrate
    rng = np.random.default_rng(0)
    m, n = 50, 120
    rows, cols = rng.integers(0, m, 600), rng.integers(0, n, 600)
    vals = rng.standard_normal(600) * 10.0 ** rng.uniform(-4, 4, 600)
    dr, dc, it = ruiz_equilibrate(m, n, rows, cols, vals, max_iter=30, tol=1e-6)
    v = np.abs(vals) / dr[rows] / dc[cols]
    r, c = np.zeros(m), np.zeros(n)
    np.maximum.at(r, rows, v)
    np.maximum.at(c, cols, v)
    assert np.abs(r[r > 0] - 1).max() < 1e-5 and np.abs(c[c > 0] - 1).max() < 1e-5 and it < 30


def test_degenerate_lp_generator_and_oracle():
    """The degenerate / rank-deficient generator: the constructed point is optimal (the oracle reaches its objective), and
    duplicated rows make the normal equations singular for the oracle as well (NormalKKTSystem needs full row rank)."""
    from madipm_jl_b200.problems import degenerate_lp
    from oracle.mpc_oracle import madipm
    qp = degenerate_lp(60, 240, 4, 21)
    r = madipm(qp, kkt_system="Normal")
    assert r.status == "SOLVE_SUCCEEDED"
    assert abs(r.objective - qp.meta["objective"]) <= 1e-6 * max(1.0, abs(r.objective))
    dup = degenerate_lp(60, 240, 4, 22, n_dup=3)
    assert dup.ncon == 63
    with np.errstate(all="ignore"):
        assert madipm(dup, kkt_system="Normal").status != "SOLVE_SUCCEEDED"
    rk = madipm(dup, kkt_system="K2")
    assert rk.status == "SOLVE_SUCCEEDED"
    assert abs(rk.objective - dup.meta["objective"]) <= 1e-6 * max(1.0, abs(rk.objective))
