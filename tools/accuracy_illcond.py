"""Accuracy of the supernodal factorization + solves on IPM-like ill-conditioned normal matrices A D A' (D spans `span`
decades, half of the columns 'basic'): normwise backward error and forward error against a dense long-double-free
reference (numpy Cholesky in float64 on the same matrix, and mpmath-free check through the residual), for Cholesky and
LDL' kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp, scipy.linalg as sl, torch
from madipm_jl_b200 import _lib
from madipm_jl_b200.problems import random_sparse_lp
from oracle import sparse_ref

def mat(m, n, k, span, seed):
    qp = random_sparse_lp(m, n, k, seed, structure="uniform")
    Bp, Bj, Bm = _lib.coo_to_csr(m, n, qp.Arows, qp.Acols)
    Cp, Cj = sparse_ref.build_normal_system(m, n, Bp, Bj)
    rng = np.random.default_rng(seed)
    d = np.where(rng.random(n) < 0.5 * m / n, 10.0 ** (span / 2), 10.0 ** (-span / 2)) * rng.uniform(0.5, 2, n)
    Cx = sparse_ref.assemble_normal_system(m, n, Bp, Bj, qp.Avals[Bm], Cp, Cj, d)
    low = sp.csc_matrix((Cx, Cj, Cp), shape=(m, m))
    return Cp, Cj, Cx, (low + sp.tril(low, -1).T).toarray()

for m, n in ((300, 1200), (1000, 4000)):
    for span in (0, 8, 12, 16):
        Cp, Cj, Cx, K = mat(m, n, 5, span, 7)
        xs = np.random.default_rng(1).standard_normal(m)
        b = K @ xs
        # float64 dense Cholesky as the yardstick
        try:
            c = sl.cho_factor(K)
            xr = sl.cho_solve(c, b)
            ref = (np.abs(K @ xr - b).max() / (np.abs(K).sum(1).max() * np.abs(xr).max()), np.linalg.norm(xr - xs) / np.linalg.norm(xs))
        except Exception as e:
            ref = ("fail", repr(e)[:40])
        out = []
        for kind in (_lib.MIPM_CHOLESKY, _lib.MIPM_LDL_DEFINITE):
            h = _lib.Handle(0)
            h.ls_analyze(m, Cp, Cj, kind=kind)
            nz = torch.tensor(Cx, device="cuda")
            ok = h.ls_factorize(nz)
            for ir in (0, 2):
                x = torch.tensor(b, device="cuda")
                h.ls_solve(x, ir)
                xg = x.cpu().numpy()
                out.append((kind, ok, ir, "%.1e" % (np.abs(K @ xg - b).max() / (np.abs(K).sum(1).max() * np.abs(xg).max())), "%.1e" % (np.linalg.norm(xg - xs) / np.linalg.norm(xs))))
        print("m", m, "span", span, "cond %.1e" % np.linalg.cond(K), "lapack (berr, ferr)", ref, "gpu (kind, ok, ir, berr, ferr)", out, flush=True)
