"""oracle/sparse_ref.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end for oracle/sparse_ref.c (the plain-C restatement of
/root/reference/src/utils.jl:158-308 and of the LDL^T algorithm behind `LDLSolver`).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this module. All public functions take and return 0-based numpy arrays.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libsparse_ref.so")
    src = os.path.join(_HERE, "sparse_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.ref_build_normal_system.restype = ctypes.c_int64
        _LIB.ref_ldl_numeric.restype = ctypes.c_int64
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _i64(v):
    return ctypes.c_int64(int(v))


def coo_to_csr(n_rows, n_cols, Ai, Aj, Ax):
    """src/utils.jl:158-201 (0-based in and out)."""
    Ai = np.ascontiguousarray(Ai, dtype=np.int32)
    Aj = np.ascontiguousarray(Aj, dtype=np.int32)
    Ax = np.ascontiguousarray(Ax, dtype=np.float64)
    nnz = len(Ai)
    Bp = np.zeros(n_rows + 1, dtype=np.int32)
    Bj = np.zeros(nnz, dtype=np.int32)
    Bx = np.zeros(nnz, dtype=np.float64)
    lib().ref_coo_to_csr(_i64(n_rows), _i64(nnz), _p(Ai), _p(Aj), _p(Ax), _p(Bp), _p(Bj), _p(Bx))
    return Bp, Bj, Bx


def build_normal_system(n_rows, n_cols, Jtp, Jtj):
    """src/utils.jl:209-274: (Cp, Cj) of tril(A A^T), column i holds rows j >= i."""
    Jtp = np.ascontiguousarray(Jtp, dtype=np.int32)
    Jtj = np.ascontiguousarray(Jtj, dtype=np.int32)
    cnt = np.zeros(n_rows + 1, dtype=np.int32)
    nnz = lib().ref_build_normal_system(_i64(n_rows), _i64(n_cols), _p(Jtp), _p(Jtj), _p(cnt), None)
    Cp = np.zeros(n_rows + 1, dtype=np.int32)
    np.cumsum(cnt[:-1], out=Cp[1:])
    Cj = np.zeros(nnz, dtype=np.int32)
    nnz2 = lib().ref_build_normal_system(_i64(n_rows), _i64(n_cols), _p(Jtp), _p(Jtj), _p(cnt), _p(Cj))
    assert nnz == nnz2 == Cp[-1]
    return Cp, Cj


def assemble_normal_system(n_rows, n_cols, Jtp, Jtj, Jtx, Cp, Cj, Dx):
    """src/utils.jl:276-308."""
    Jtp = np.ascontiguousarray(Jtp, dtype=np.int32)
    Jtj = np.ascontiguousarray(Jtj, dtype=np.int32)
    Jtx = np.ascontiguousarray(Jtx, dtype=np.float64)
    Cp = np.ascontiguousarray(Cp, dtype=np.int32)
    Cj = np.ascontiguousarray(Cj, dtype=np.int32)
    Dx = np.ascontiguousarray(Dx, dtype=np.float64)
    Cx = np.zeros(len(Cj), dtype=np.float64)
    lib().ref_assemble_normal_system(_i64(n_rows), _i64(n_cols), _p(Jtp), _p(Jtj), _p(Jtx),
                                     _p(Cp), _p(Cj), _p(Cx), _p(Dx))
    return Cx


def transfer(nnz_dest, src, map_):
    """MadNLP.transfer!: zero dest then dest[map[k]] += src[k] (k ascending)."""
    src = np.ascontiguousarray(src, dtype=np.float64)
    map_ = np.ascontiguousarray(map_, dtype=np.int64)
    dest = np.zeros(nnz_dest, dtype=np.float64)
    lib().ref_transfer(_i64(nnz_dest), _p(dest), _i64(len(src)), _p(src), _p(map_))
    return dest


class LDL:
    """Up-looking sparse LDL^T (Davis 2005) = the algorithm of LDLFactorizations.jl, the
    `LDLSolver` the reference's tests select (test/runtests.jl:128,185).

    Built from the LOWER-triangular CSC of a symmetric matrix (what MadNLP hands every
    linear solver: SURVEY 8b); lower CSC == upper CSR, so we transpose once to the upper
    CSC the algorithm wants. `perm` is a fill-reducing ordering (new -> old) or None.
    """

    def __init__(self, n, colptr, rowval, perm=None):
        import scipy.sparse as sp
        self.n = n
        low = sp.csc_matrix((np.arange(1, len(rowval) + 1, dtype=np.float64), rowval, colptr), shape=(n, n))
        up = low.T.tocsc()  # upper triangular CSC, data = 1-based index into the lower nzval
        if perm is not None:
            # with a permutation the algorithm reads the upper triangle of P A P', whose entries
            # come from either triangle of A: hand it the full symmetric pattern
            up = (up + sp.tril(low, -1)).tocsc()
        up.sort_indices()
        self.Ap = up.indptr.astype(np.int64)
        self.Ai = up.indices.astype(np.int32)
        self.src = (up.data - 1).astype(np.int64)
        if perm is not None:
            self.P = np.ascontiguousarray(perm, dtype=np.int32)
            self.Pinv = np.empty(n, dtype=np.int32)
            self.Pinv[self.P] = np.arange(n, dtype=np.int32)
        else:
            self.P = self.Pinv = None
        self.Lp = np.zeros(n + 1, dtype=np.int64)
        self.Parent = np.zeros(n, dtype=np.int32)
        self.Lnz = np.zeros(n, dtype=np.int64)
        self.Flag = np.zeros(n, dtype=np.int32)
        lib().ref_ldl_symbolic(_i64(n), _p(self.Ap), _p(self.Ai), _p(self.Lp), _p(self.Parent),
                               _p(self.Lnz), _p(self.Flag), _p(self.P), _p(self.Pinv))
        self.colcount = self.Lnz.copy()
        self.nnzL = int(self.Lp[-1])
        self.flops = float(np.sum((self.colcount.astype(np.float64) + 1.0) ** 2))
        self.Li = np.zeros(max(self.nnzL, 1), dtype=np.int32)
        self.Lx = np.zeros(max(self.nnzL, 1), dtype=np.float64)
        self.D = np.zeros(n, dtype=np.float64)
        self.Y = np.zeros(n, dtype=np.float64)
        self.Pattern = np.zeros(n, dtype=np.int32)
        self.W = np.zeros(n, dtype=np.float64)
        self.ok = False

    def factorize(self, nzval):
        Ax = np.ascontiguousarray(np.asarray(nzval, dtype=np.float64)[self.src])
        d = lib().ref_ldl_numeric(_i64(self.n), _p(self.Ap), _p(self.Ai), _p(Ax), _p(self.Lp),
                                  _p(self.Parent), _p(self.Lnz), _p(self.Li), _p(self.Lx), _p(self.D),
                                  _p(self.Y), _p(self.Pattern), _p(self.Flag), _p(self.P), _p(self.Pinv))
        self.ok = (d == self.n) and bool(np.all(np.isfinite(self.D)))
        return self.ok

    def solve(self, b):
        x = np.array(b, dtype=np.float64, copy=True)
        lib().ref_ldl_solve(_i64(self.n), _p(x), _p(self.Lp), _p(self.Li), _p(self.Lx), _p(self.D),
                            _p(self.P), _p(self.W))
        return x

    def inertia(self):
        return int(np.sum(self.D > 0)), int(np.sum(self.D == 0)), int(np.sum(self.D < 0))
