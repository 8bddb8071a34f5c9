"""CPU tests of the C-ABI library: it loads, exports exactly what include/madipm_b200.h declares,
and its HOST symbolic entry points are bit-exact with the oracle's literal restatement.
No compute entry point is exercised here (they need a GPU and refuse an analysis-only handle)."""
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from madipm_jl_b200 import _lib
from madipm_jl_b200.problems import random_sparse_lp, random_sparse_qp
from oracle import sparse_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "madipm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mipm_[a-z0-9_]+)\s*\(", hdr)))
    lib = _lib.load()
    assert declared, "no declarations parsed"
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == declared
    assert lib.mipm_version() >= 100


def test_analysis_only_handle_refuses_device_work(built):
    h = _lib.Handle(device=-1)
    with pytest.raises(_lib.MipmError) as e:
        h.set_predictive_rhs()
    assert e.value.code == _lib.MIPM_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(_lib.MipmError):
        h.ls_solve(np.zeros(3), 0)


@pytest.mark.parametrize("case", [(1, 2, 2, "uniform"), (60, 240, 5, "uniform"), (400, 2000, 5, "window"),
                                  (1500, 6000, 6, "window")])
def test_coo_to_csr_and_normal_pattern_bit_exact(built, case):
    m, n, k, structure = case
    qp = random_sparse_lp(m, n, min(k, m), 3, structure=structure, window=15)
    Bp, Bj, Bm = _lib.coo_to_csr(m, n, qp.Arows, qp.Acols)
    rp, rj, rx = sparse_ref.coo_to_csr(m, n, qp.Arows, qp.Acols, np.arange(len(qp.Arows), dtype=float))
    assert (Bp == rp).all() and (Bj == rj).all() and (Bm == rx.astype(np.int64)).all()
    Cp, Cj = _lib.Handle(device=-1).normal_symbolic(m, n, Bp, Bj)
    Rp, Rj = sparse_ref.build_normal_system(m, n, rp, rj)
    assert Cp.dtype == np.int32 and (Cp == Rp).all() and (Cj == Rj).all()
    # the Julia convention: 1-based Int32 in and out
    Cp1, Cj1 = _lib.Handle(device=-1).normal_symbolic(m, n, Bp + 1, Bj + 1, index_base=1)
    assert (Cp1 == Rp + 1).all() and (Cj1 == Rj + 1).all()


def test_normal_symbolic_edge_cases(built):
    h = _lib.Handle(device=-1)
    # empty rows: the diagonal is present iff the row is non-empty
    Ap = np.array([0, 2, 2, 3], dtype=np.int32)
    Aj = np.array([0, 1, 1], dtype=np.int32)
    Cp, Cj = h.normal_symbolic(3, 2, Ap, Aj)
    Rp, Rj = sparse_ref.build_normal_system(3, 2, Ap, Aj)
    assert (Cp == Rp).all() and (Cj == Rj).all()
    assert list(Cp) == [0, 2, 2, 3] and list(Cj) == [0, 2, 2]
    # zero-size
    Cp, Cj = h.normal_symbolic(0, 0, np.array([0], dtype=np.int32), np.zeros(0, dtype=np.int32))
    assert list(Cp) == [0] and len(Cj) == 0
    # duplicate column inside a row is rejected (the reference's buffer[k] overwrite is ill-defined)
    with pytest.raises(_lib.MipmError) as e:
        h.normal_symbolic(1, 2, np.array([0, 2], dtype=np.int32), np.array([1, 1], dtype=np.int32))
    assert e.value.code == _lib.MIPM_ERR_DUPLICATE
    with pytest.raises(_lib.MipmError):
        h.normal_symbolic(1, 2, np.array([0, 1], dtype=np.int32), np.array([5], dtype=np.int32))


def test_k2_symbolic_matches_sparse_pattern(built):
    qp = random_sparse_qp(50, 120, 4, 2, structure="window", window=10)
    n, m = qp.nvar, qp.ncon
    I = np.concatenate([np.arange(n), qp.Hrows, n + qp.Arows, n + np.arange(m)]).astype(np.int32)
    J = np.concatenate([np.arange(n), qp.Hcols, qp.Acols, n + np.arange(m)]).astype(np.int32)
    colptr, rowval, kmap = _lib.Handle(device=-1).k2_symbolic(n + m, I, J)
    V = np.random.default_rng(0).standard_normal(len(I))
    ref = sp.csc_matrix((V, (I, J)), shape=(n + m, n + m))   # sums duplicates, sorts rows
    ref.sort_indices()
    assert (ref.indptr == colptr).all() and (ref.indices == rowval).all()
    assert np.allclose(sparse_ref.transfer(len(rowval), V, kmap), ref.data, atol=1e-14)
    # diagonal of Q and pr_diag share a slot: the map has duplicates
    assert len(np.unique(kmap)) == len(rowval) < len(kmap)
    with pytest.raises(_lib.MipmError):
        _lib.Handle(device=-1).k2_symbolic(3, np.array([0], dtype=np.int32), np.array([2], dtype=np.int32))


def _multifrontal_numpy(n, colptr, rowval, vals, sym, ldl=False):
    """Dense multifrontal factorization driven ONLY by the library's symbolic structure."""
    perm, snp, spar, rp, ri = sym["perm"], sym["sn_ptr"], sym["sn_parent"], sym["row_ptr"], sym["row_idx"]
    A = sp.csc_matrix((vals, rowval, colptr), shape=(n, n))
    A = (A + sp.tril(A, -1).T).toarray()[np.ix_(perm, perm)]
    ns = len(snp) - 1
    kids = [[] for _ in range(ns)]
    for s in range(ns):
        if spar[s] >= 0:
            assert spar[s] > s
            kids[spar[s]].append(s)
    U = [None] * ns
    L = np.zeros((n, n))
    D = np.ones(n)
    for s in range(ns):
        c0, c1 = snp[s], snp[s + 1]
        rows = ri[rp[s]:rp[s + 1]]
        assert np.all(np.diff(rows) > 0) and (len(rows) == 0 or rows[0] >= c1)
        idx = np.concatenate([np.arange(c0, c1), rows])
        k = c1 - c0
        F = np.zeros((len(idx), len(idx)))
        F[:, :k] = A[np.ix_(idx, np.arange(c0, c1))]
        F[:k, :] = F[:, :k].T
        pos = {g: t for t, g in enumerate(idx)}
        for c in kids[s]:
            rel = np.array([pos[g] for g in ri[rp[c]:rp[c + 1]]], dtype=int)
            F[np.ix_(rel, rel)] += U[c]
        if not ldl:
            L11 = np.linalg.cholesky(F[:k, :k])
            L21 = np.linalg.solve(L11, F[k:, :k].T).T
            U[s] = F[k:, k:] - L21 @ L21.T
        else:
            L11 = np.eye(k)
            d = np.zeros(k)
            W = F[:k, :k].copy()
            for j in range(k):
                d[j] = W[j, j]
                L11[j + 1:, j] = W[j + 1:, j] / d[j]
                W[j + 1:, j + 1:] -= np.outer(L11[j + 1:, j], L11[j + 1:, j]) * d[j]
            X = np.linalg.solve(L11, F[k:, :k].T).T
            L21 = X / d
            U[s] = F[k:, k:] - X @ L21.T
            D[c0:c1] = d
        L[c0:c1, c0:c1] = L11
        L[rows, c0:c1] = L21
    return A, L, D


@pytest.mark.parametrize("case", [(60, 240, 5, "uniform"), (400, 2000, 5, "window"), (900, 3000, 4, "window")])
def test_ls_symbolic_structure_drives_a_correct_factorization(built, case):
    m, n, k, structure = case
    qp = random_sparse_lp(m, n, k, 3, structure=structure, window=15)
    Bp, Bj, Bm = _lib.coo_to_csr(m, n, qp.Arows, qp.Acols)
    h = _lib.Handle(device=-1)
    Cp, Cj = h.normal_symbolic(m, n, Bp, Bj)
    h.ls_analyze(m, Cp, Cj)
    st, sym = h.ls_stats(), h.ls_symbolic()
    assert sorted(sym["perm"]) == list(range(m))
    assert st["nnz_l"] >= st["nnz_l_exact"] > 0 and st["n_supernodes"] == len(sym["sn_ptr"]) - 1
    D = np.random.default_rng(0).uniform(0.5, 2.0, n)
    Cx = sparse_ref.assemble_normal_system(m, n, Bp, Bj, qp.Avals[Bm], Cp, Cj, D)
    A, L, _ = _multifrontal_numpy(m, Cp, Cj, Cx, sym)
    assert np.abs(L @ L.T - A).max() < 1e-10 * np.abs(A).max()
    # deterministic analysis: same input -> identical structure (the "symbolic structure bit-exact" gate)
    h2 = _lib.Handle(device=-1)
    h2.normal_symbolic(m, n, Bp, Bj)
    h2.ls_analyze(m, Cp, Cj)
    sym2 = h2.ls_symbolic()
    assert all((sym[key] == sym2[key]).all() for key in sym)
    # orderings: natural and user-supplied
    h3 = _lib.Handle(device=-1)
    h3.ls_analyze(m, Cp, Cj, ordering=_lib.MIPM_ORDER_USER, user_perm=np.arange(m)[::-1].copy())
    A3, L3, _ = _multifrontal_numpy(m, Cp, Cj, Cx, h3.ls_symbolic())
    assert np.abs(L3 @ L3.T - A3).max() < 1e-10 * np.abs(A).max()


def test_ls_symbolic_k2_ldl_ordering_is_quasidefinite_safe(built):
    """For K2 = [Q+S A'; A dI] with the reference's default delta_c = +1e-10 (quirk A.9 v, not
    quasi-definite) the static-pivot LDL^T must stay stable: checked through the ordering property
    and a growth-free reconstruction with Sigma spanning ten orders of magnitude."""
    qp = random_sparse_qp(50, 120, 4, 2, structure="window", window=10)
    n, m = qp.nvar, qp.ncon
    I = np.concatenate([np.arange(n), qp.Hrows, n + qp.Arows, n + np.arange(m)]).astype(np.int32)
    J = np.concatenate([np.arange(n), qp.Hcols, qp.Acols, n + np.arange(m)]).astype(np.int32)
    h = _lib.Handle(device=-1)
    colptr, rowval, kmap = h.k2_symbolic(n + m, I, J)
    h.ls_analyze(n + m, colptr, rowval, kind=_lib.MIPM_LDL)
    sym = h.ls_symbolic()
    iperm = np.empty(n + m, dtype=int)
    iperm[sym["perm"]] = np.arange(n + m)
    A = sp.csr_matrix((qp.Avals, (qp.Arows, qp.Acols)), shape=(m, n))
    for i in range(m):
        cols = A.indices[A.indptr[i]:A.indptr[i + 1]]
        assert iperm[n + i] > iperm[cols].min()
    # every dual vertex comes after ALL of its primal neighbours -> its pivot is the (strongly negative)
    # Schur complement delta_c - sum a^2/piv, never the bare delta_c: no growth even with badly scaled Sigma
    for i in range(m):
        cols = A.indices[A.indptr[i]:A.indptr[i + 1]]
        assert iperm[n + i] > iperm[cols].max()
    rng = np.random.default_rng(1)
    V = np.concatenate([10.0 ** rng.uniform(-6, 4, n), qp.Hvals, qp.Avals, np.full(m, 1e-10)])
    nz = sparse_ref.transfer(len(rowval), V, kmap)
    Ad, L, D = _multifrontal_numpy(n + m, colptr, rowval, nz, sym, ldl=True)
    assert np.abs(L @ np.diag(D) @ L.T - Ad).max() < 1e-12 * np.abs(Ad).max()
    assert np.abs(L).max() < 1e3
    assert (D > 0).sum() == n and (D < 0).sum() == m


def test_host_analysis_is_independent_of_the_thread_count(built, monkeypatch):
    """The pattern / product-term sweep and the nested dissection run on a pool of host threads (MIPM_HOST_THREADS);
    the symbolic structure must be identical for any thread count. Sized so that both really use several threads
    (row chunks of 4096 rows, one dissection worker per 20000 vertices)."""
    from madipm_jl_b200.problems import random_sparse_lp
    m, n = 50_000, 150_000
    qp = random_sparse_lp(m, n, 4, 21, structure="window", window=40)
    Bp, Bj, _ = _lib.coo_to_csr(m, n, qp.Arows, qp.Acols)
    got = {}
    for nt in (1, 5):
        monkeypatch.setenv("MIPM_HOST_THREADS", str(nt))
        h = _lib.Handle(device=-1)
        Cp, Cj = h.normal_symbolic(m, n, Bp, Bj)
        h.ls_analyze(m, Cp, Cj)
        got[nt] = (Cp, Cj, h.ls_symbolic(), h.ls_stats())
        h.close()
    a, b = got[1], got[5]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    for key in a[2]:
        assert np.array_equal(a[2][key], b[2][key]), key
    assert a[3] == b[3]


def test_ls_analyze_edge_cases(built):
    """Host analysis on degenerate inputs: empty and 1 x 1 matrices, diagonal-only patterns, missing diagonals with 1-based
    indices, and the input checks (non-monotone column pointer, entry above the diagonal, pointer not starting at the base)."""
    h = _lib.Handle(device=-1)
    h.ls_analyze(0, np.array([0], dtype=np.int32), np.zeros(0, dtype=np.int32))
    assert h.ls_stats()["n_supernodes"] == 0
    h.ls_analyze(1, np.array([0, 1], dtype=np.int32), np.array([0], dtype=np.int32))
    assert h.ls_stats()["n"] == 1
    n = 5000
    h.ls_analyze(n, np.arange(n + 1, dtype=np.int32), np.arange(n, dtype=np.int32))
    assert h.ls_stats()["nnz_l_exact"] == n
    h.ls_analyze(3, np.array([1, 2, 3, 3], dtype=np.int32), np.array([2, 3], dtype=np.int32), index_base=1)
    assert h.ls_stats()["nnz_l_exact"] == 5                  # chain 1 - 2 - 3: no fill, diagonals counted
    for cp, ri in (([0, 2, 1], [0, 1]), ([0, 1, 2], [1, 0]), ([1, 2, 3], [0, 1])):
        with pytest.raises(_lib.MipmError):
            h.ls_analyze(2, np.array(cp, dtype=np.int32), np.array(ri, dtype=np.int32))
    Bp, Bj, Bm = _lib.coo_to_csr(3, 2, np.zeros(0, dtype=np.int32), np.zeros(0, dtype=np.int32))
    assert list(Bp) == [0, 0, 0, 0] and len(Bj) == 0
    for I, J in (([3], [0]), ([1], [2])):
        with pytest.raises(_lib.MipmError):
            _lib.coo_to_csr(3, 2, np.array(I, dtype=np.int32), np.array(J, dtype=np.int32))
