"""ctypes binding of the C ABI declared in include/madipm_b200.h.

This is the same boundary the Julia glue (ext/MadIPMB200Ext) binds with `ccall`; nothing here
computes. If libmadipm_b200.so is missing the import fails loudly -- there is no fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MIPM_LIB") or os.path.join(_HERE, "libmadipm_b200.so")      # MIPM_LIB: A/B builds (tools/)

MIPM_OK, MIPM_ERR_ARG, MIPM_ERR_CUDA, MIPM_ERR_ALLOC, MIPM_ERR_STATE, MIPM_ERR_DUPLICATE, MIPM_ERR_NOT_FACTORIZED = range(7)
MIPM_CHOLESKY, MIPM_LDL, MIPM_LDL_DEFINITE = 0, 1, 2
MIPM_ORDER_ND, MIPM_ORDER_NATURAL, MIPM_ORDER_USER = 0, 1, 2

_ERRNAMES = {1: "MIPM_ERR_ARG", 2: "MIPM_ERR_CUDA", 3: "MIPM_ERR_ALLOC", 4: "MIPM_ERR_STATE",
             5: "MIPM_ERR_DUPLICATE", 6: "MIPM_ERR_NOT_FACTORIZED"}


class MipmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_ERRNAMES.get(code, code)}: {msg}")
        self.code = code


class LsStats(C.Structure):
    _fields_ = [("n", C.c_int64), ("nnz_a", C.c_int64), ("nnz_l", C.c_int64), ("nnz_l_exact", C.c_int64),
                ("flops", C.c_double), ("n_supernodes", C.c_int64), ("n_levels", C.c_int64),
                ("max_front_cols", C.c_int64), ("max_front_rows", C.c_int64), ("update_doubles", C.c_int64),
                ("n_launches", C.c_int64)]


class MpcModel(C.Structure):
    _fields_ = [("kkt_kind", C.c_int), ("exact_order", C.c_int), ("nx", C.c_int64), ("c0", C.c_double),
                ("d_ATx", C.c_void_p), ("d_cvec", C.c_void_p), ("d_Hx", C.c_void_p), ("d_aug_nz", C.c_void_p),
                ("d_aug_raw_V", C.c_void_p), ("d_buffer_n", C.c_void_p), ("d_buffer_m", C.c_void_p)]


class MpcVectors(C.Structure):
    _fields_ = [("n", C.c_int64), ("m", C.c_int64), ("nlb", C.c_int64), ("nub", C.c_int64),
                ("index_base", C.c_int),
                ("d_ind_lb", C.c_void_p), ("d_ind_ub", C.c_void_p),
                ("d_x", C.c_void_p), ("d_xl", C.c_void_p), ("d_xu", C.c_void_p), ("d_zl", C.c_void_p),
                ("d_zu", C.c_void_p), ("d_f", C.c_void_p),
                ("d_y", C.c_void_p), ("d_c", C.c_void_p), ("d_rhs", C.c_void_p),
                ("d_jacl", C.c_void_p),
                ("d_d", C.c_void_p), ("d_p", C.c_void_p), ("d_w", C.c_void_p),
                ("d_corr_lb", C.c_void_p), ("d_corr_ub", C.c_void_p),
                ("d_reg", C.c_void_p), ("d_pr_diag", C.c_void_p), ("d_du_diag", C.c_void_p),
                ("d_l_diag", C.c_void_p), ("d_u_diag", C.c_void_p), ("d_l_lower", C.c_void_p),
                ("d_u_lower", C.c_void_p)]


# every symbol include/madipm_b200.h declares (tests check the .so exports exactly these)
SYMBOLS = [
    "mipm_version", "mipm_create", "mipm_destroy", "mipm_last_error", "mipm_free",
    "mipm_coo_to_csr", "mipm_normal_symbolic", "mipm_normal_set_jacobian", "mipm_normal_assemble",
    "mipm_k2_symbolic", "mipm_k2_transfer",
    "mipm_ls_analyze", "mipm_ls_factorize", "mipm_ls_factorize_async", "mipm_ls_status", "mipm_ls_solve",
    "mipm_ls_inertia", "mipm_ls_stats", "mipm_ls_symbolic",
    "mipm_ls_analyze_border", "mipm_ls_factorize_stage", "mipm_ls_solve_stage", "mipm_ls_root_info",
    "mipm_set_grid_limit", "mipm_spmv_setup", "mipm_spmv", "mipm_spmv_pair", "mipm_spmv_cache_values", "mipm_hess_setup", "mipm_hess_spmv",
    "mipm_mpc_set_model", "mipm_mpc_iter_begin", "mipm_mpc_peek", "mipm_mpc_refactor", "mipm_mpc_iter_rest",
    "mipm_mpc_bind", "mipm_set_aug_diagonal_reg", "mipm_set_predictive_rhs", "mipm_set_correction_rhs",
    "mipm_get_correction", "mipm_set_extra_correction", "mipm_get_complementarity_measure",
    "mipm_get_affine_complementarity_measure", "mipm_get_alpha_max", "mipm_termination_measures",
    "mipm_apply_step", "mipm_reduce_rhs", "mipm_finish_aug_solve", "mipm_normal_solve_stage", "mipm_kktmul",
    "mipm_residual_norms", "mipm_init_point_stage", "mipm_init_bounds", "mipm_amax", "mipm_axpby", "mipm_fill", "mipm_copy", "mipm_gather", "mipm_scatter", "mipm_dot",
    "mipm_launch_count", "mipm_ls_factorize_profile", "mipm_bench_syrk", "mipm_mehrotra_adaptive_step", "mipm_set_aug_diagonal_reg_scaled", "mipm_k25_scale_values",
    "mipm_reduce_rhs_scaled", "mipm_finish_aug_solve_scaled", "mipm_kktmul_scaled",
    "mipm_batch_configure", "mipm_batch_set_active", "mipm_batch_amax", "mipm_batch_dot", "mipm_batch_init_point_stage",
    "mipm_batch_iter_begin", "mipm_batch_peek", "mipm_batch_iter_rest", "mipm_ruiz_equilibrate", "mipm_scale_coo", "mipm_mpc_ext_begin", "mipm_mpc_ext_fetch", "mipm_mpc_ext_phase",
]

_lib = None


def load():
    """Load libmadipm_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C madipm_jl_b200/csrc`). The CUDA library is mandatory; there is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.mipm_last_error.restype = C.c_char_p
    lib.mipm_last_error.argtypes = [C.c_void_p]
    lib.mipm_launch_count.restype = C.c_int64
    lib.mipm_launch_count.argtypes = [C.c_void_p]
    lib.mipm_free.restype = None
    lib.mipm_free.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _ptr(a):
    """Device pointer of a torch tensor, host pointer of a numpy array, or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    if isinstance(a, int):
        return C.c_void_p(a)
    raise TypeError(type(a))


class _Owner:
    """Frees a library-owned host array when the last numpy view of it is gone."""

    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        try:
            self.lib.mipm_free(C.c_void_p(self.ptr))
        except Exception:
            pass


def _take(ptr, count, ctype, dtype):
    """numpy view of a library-owned host array (no copy: the patterns the symbolic entry points return are tens of
    megabytes). The ctypes buffer every view ends up referencing carries the owner object that calls mipm_free."""
    lib = load()
    if count <= 0 or np.dtype(ctype) != np.dtype(dtype):
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(max(count, 1),))[:max(count, 0)].astype(dtype, copy=True)
        lib.mipm_free(ptr)
        return arr
    buf = (ctype * count).from_address(ptr.value)
    buf._owner = _Owner(lib, ptr.value)
    return np.frombuffer(buf, dtype=dtype)


class Handle:
    """RAII wrapper of mipm_handle. device=-1 gives an analysis-only (host symbolic) handle."""

    def __init__(self, device=0, stream=0):
        self.lib = load()
        self.h = C.c_void_p()
        rc = self.lib.mipm_create(C.byref(self.h), C.c_int(device), C.c_void_p(stream))
        if rc != MIPM_OK:
            raise MipmError(rc, "mipm_create failed (no CUDA device? the library has no CPU fallback)")
        self.device = device

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.mipm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != MIPM_OK:
            raise MipmError(rc, self.lib.mipm_last_error(self.h).decode())

    # ---- host symbolic
    def normal_symbolic(self, m, n, Ap=None, Aj=None, index_base=0):
        """Ap = Aj = None: the matrix registered with spmv_setup on this handle (device handles only)."""
        if Ap is not None:
            Ap = np.ascontiguousarray(Ap, dtype=np.int32)
            Aj = np.ascontiguousarray(Aj, dtype=np.int32)
        cp, cj, nnz = C.c_void_p(), C.c_void_p(), C.c_int64()
        self.check(self.lib.mipm_normal_symbolic(self.h, C.c_int64(m), C.c_int64(n), _ptr(Ap), _ptr(Aj),
                                                 C.c_int(index_base), C.byref(cp), C.byref(cj), C.byref(nnz)))
        return _take(cp, m + 1, C.c_int32, np.int32), _take(cj, nnz.value, C.c_int32, np.int32)

    def k2_symbolic(self, dim, I, J, index_base=0):
        I = np.ascontiguousarray(I, dtype=np.int32)
        J = np.ascontiguousarray(J, dtype=np.int32)
        cp, rv, mp, nnz = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
        self.check(self.lib.mipm_k2_symbolic(self.h, C.c_int64(dim), C.c_int64(len(I)), _ptr(I), _ptr(J),
                                             C.c_int(index_base), C.byref(cp), C.byref(rv), C.byref(mp), C.byref(nnz)))
        return (_take(cp, dim + 1, C.c_int32, np.int32), _take(rv, nnz.value, C.c_int32, np.int32),
                _take(mp, len(I), C.c_int64, np.int64))

    def ls_analyze(self, n, colptr, rowval, kind=MIPM_CHOLESKY, ordering=MIPM_ORDER_ND, user_perm=None, index_base=0):
        colptr = np.ascontiguousarray(colptr, dtype=np.int32)
        rowval = np.ascontiguousarray(rowval, dtype=np.int32)
        up = None if user_perm is None else np.ascontiguousarray(user_perm, dtype=np.int32)
        self.check(self.lib.mipm_ls_analyze(self.h, C.c_int64(n), _ptr(colptr), _ptr(rowval), C.c_int(index_base),
                                            C.c_int(kind), C.c_int(ordering), _ptr(up)))

    def ls_analyze_border(self, n, colptr, rowval, n_border, kind=MIPM_CHOLESKY, index_base=0):
        colptr = np.ascontiguousarray(colptr, dtype=np.int32)
        rowval = np.ascontiguousarray(rowval, dtype=np.int32)
        self.check(self.lib.mipm_ls_analyze_border(self.h, C.c_int64(n), _ptr(colptr), _ptr(rowval), C.c_int(index_base),
                                                   C.c_int(kind), C.c_int64(n_border)))

    def ls_factorize_stage(self, nzval, stage):
        if stage == 0:
            self._nz_ref = nzval
        self.check(self.lib.mipm_ls_factorize_stage(self.h, _ptr(nzval), C.c_int(stage)))

    def ls_solve_stage(self, x, stage):
        self.check(self.lib.mipm_ls_solve_stage(self.h, _ptr(x), C.c_int(stage)))

    def ls_root_info(self):
        """(device pointer of the root panel, n_root, device pointer of the root RHS segment)."""
        pp, pr, nr = C.c_void_p(), C.c_void_p(), C.c_int64()
        self.check(self.lib.mipm_ls_root_info(self.h, C.byref(pp), C.byref(nr), C.byref(pr)))
        return pp.value, nr.value, pr.value

    def ls_stats(self):
        st = LsStats()
        self.check(self.lib.mipm_ls_stats(self.h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in LsStats._fields_}

    def ls_symbolic(self):
        perm, snp, spar, rp, ri = (C.c_void_p() for _ in range(5))
        ns = C.c_int64()
        self.check(self.lib.mipm_ls_symbolic(self.h, C.byref(perm), C.byref(ns), C.byref(snp), C.byref(spar),
                                             C.byref(rp), C.byref(ri)))
        n = self.ls_stats()["n"]
        ns = ns.value
        row_ptr = _take(rp, ns + 1, C.c_int64, np.int64)
        return dict(perm=_take(perm, n, C.c_int32, np.int32), sn_ptr=_take(snp, ns + 1, C.c_int32, np.int32),
                    sn_parent=_take(spar, ns, C.c_int32, np.int32), row_ptr=row_ptr,
                    row_idx=_take(ri, int(row_ptr[-1]), C.c_int32, np.int32))

    # ---- device entry points (arguments are torch CUDA tensors)
    def normal_set_jacobian(self, ATx):
        self.check(self.lib.mipm_normal_set_jacobian(self.h, _ptr(ATx)))

    def normal_assemble(self, pr_diag, Cx, exact_order=False):
        self.check(self.lib.mipm_normal_assemble(self.h, _ptr(pr_diag), _ptr(Cx), C.c_int(int(exact_order))))

    def k2_transfer(self, V, nz):
        self.check(self.lib.mipm_k2_transfer(self.h, _ptr(V), _ptr(nz)))

    def ls_factorize(self, nzval):
        self._nz_ref = nzval      # refinement reads these values later: keep the buffer alive
        st = C.c_int()
        self.check(self.lib.mipm_ls_factorize(self.h, _ptr(nzval), C.byref(st)))
        return st.value == MIPM_OK

    def ls_factorize_async(self, nzval):
        self._nz_ref = nzval
        self.check(self.lib.mipm_ls_factorize_async(self.h, _ptr(nzval)))

    def ls_status(self):
        st = C.c_int()
        self.check(self.lib.mipm_ls_status(self.h, C.byref(st)))
        return st.value == MIPM_OK

    def ls_solve(self, x, ir_steps=0):
        self.check(self.lib.mipm_ls_solve(self.h, _ptr(x), C.c_int(ir_steps)))

    def ls_inertia(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self.check(self.lib.mipm_ls_inertia(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def spmv_setup(self, m, n, Ap, Aj, index_base=0):
        Ap = np.ascontiguousarray(Ap, dtype=np.int32)
        Aj = np.ascontiguousarray(Aj, dtype=np.int32)
        self.check(self.lib.mipm_spmv_setup(self.h, C.c_int64(m), C.c_int64(n), _ptr(Ap), _ptr(Aj), C.c_int(index_base)))

    def spmv(self, trans, alpha, Ax, x, beta, y):
        self.check(self.lib.mipm_spmv(self.h, C.c_int(trans), C.c_double(alpha), _ptr(Ax), _ptr(x), C.c_double(beta), _ptr(y)))

    def spmv_pair(self, Ax, alpha1, x1, beta1, y1, alpha2, x2, beta2, y2):
        """y1 = alpha1 A x1 + beta1 y1 and y2 = alpha2 A' x2 + beta2 y2 in one launch."""
        self.check(self.lib.mipm_spmv_pair(self.h, _ptr(Ax), C.c_double(alpha1), _ptr(x1), C.c_double(beta1), _ptr(y1),
                                           C.c_double(alpha2), _ptr(x2), C.c_double(beta2), _ptr(y2)))

    def set_grid_limit(self, max_ctas):
        self.check(self.lib.mipm_set_grid_limit(self.h, C.c_int(int(max_ctas))))

    def spmv_cache_values(self, Ax):
        self.check(self.lib.mipm_spmv_cache_values(self.h, _ptr(Ax) if Ax is not None else None))

    def hess_setup(self, n, Hp, Hj, index_base=0):
        Hp = np.ascontiguousarray(Hp, dtype=np.int32)
        Hj = np.ascontiguousarray(Hj, dtype=np.int32)
        self.check(self.lib.mipm_hess_setup(self.h, C.c_int64(n), _ptr(Hp), _ptr(Hj), C.c_int(index_base)))

    def hess_spmv(self, alpha, Hx, x, beta, y):
        self.check(self.lib.mipm_hess_spmv(self.h, C.c_double(alpha), _ptr(Hx), _ptr(x), C.c_double(beta), _ptr(y)))

    def mpc_set_model(self, model: "MpcModel"):
        self.check(self.lib.mipm_mpc_set_model(self.h, C.byref(model)))

    def mpc_iter_begin(self, del_w, del_c):
        out = (C.c_double * 16)()
        st = C.c_int()
        self.check(self.lib.mipm_mpc_iter_begin(self.h, C.c_double(del_w), C.c_double(del_c), out, C.byref(st)))
        return list(out), st.value == MIPM_OK

    def mpc_peek(self):
        out = (C.c_double * 16)()
        self.check(self.lib.mipm_mpc_peek(self.h, out))
        return list(out)

    def mpc_refactor(self, del_w, del_c):
        st = C.c_int()
        self.check(self.lib.mipm_mpc_refactor(self.h, C.c_double(del_w), C.c_double(del_c), C.byref(st)))
        return st.value == MIPM_OK

    def mpc_iter_rest(self, mu_min, step_rule, tau_param, ir_steps):
        self.check(self.lib.mipm_mpc_iter_rest(self.h, C.c_double(mu_min), C.c_int(step_rule), C.c_double(tau_param), C.c_int(ir_steps)))

    # ---- fused iteration around an external linear solver
    def mpc_ext_begin(self, del_w, del_c):
        self.check(self.lib.mipm_mpc_ext_begin(self.h, C.c_double(del_w), C.c_double(del_c)))

    def mpc_ext_fetch(self):
        out = (C.c_double * 16)()
        self.check(self.lib.mipm_mpc_ext_fetch(self.h, out))
        return list(out)

    def mpc_ext_phase(self, phase, mu_min, step_rule, tau_param):
        self.check(self.lib.mipm_mpc_ext_phase(self.h, C.c_int(phase), C.c_double(mu_min), C.c_int(step_rule), C.c_double(tau_param)))

    # ---- preprocessing
    def ruiz_equilibrate(self, m, n, rows, cols, vals, Dr, Dc, max_iter=10, tol=0.0, index_base=0):
        it = C.c_int()
        self.check(self.lib.mipm_ruiz_equilibrate(self.h, C.c_int64(m), C.c_int64(n), C.c_int64(vals.numel()), _ptr(rows), _ptr(cols),
                                                  _ptr(vals), C.c_int(index_base), C.c_int(max_iter), C.c_double(tol), _ptr(Dr), _ptr(Dc),
                                                  C.byref(it)))
        return it.value

    def scale_coo(self, rows, cols, vals, Dr, Dc, out, index_base=0):
        self.check(self.lib.mipm_scale_coo(self.h, C.c_int64(vals.numel()), _ptr(rows), _ptr(cols), _ptr(vals), C.c_int(index_base),
                                           _ptr(Dr), _ptr(Dc), _ptr(out)))

    # ---- batches of stacked independent units (BASELINE config C5)
    def batch_configure(self, off_n, off_m):
        off_n = np.ascontiguousarray(off_n, dtype=np.int64)
        off_m = np.ascontiguousarray(off_m, dtype=np.int64)
        self._nb = len(off_n) - 1
        self.check(self.lib.mipm_batch_configure(self.h, C.c_int64(self._nb), _ptr(off_n), _ptr(off_m)))

    def batch_set_active(self, active):
        a = np.ascontiguousarray(active, dtype=np.int32)
        self.check(self.lib.mipm_batch_set_active(self.h, _ptr(a)))

    def batch_amax(self, by_rows, x):
        out = np.zeros(self._nb)
        self.check(self.lib.mipm_batch_amax(self.h, C.c_int(int(by_rows)), _ptr(x), _ptr(out)))
        return out

    def batch_dot(self, x, y):
        out = np.zeros(self._nb)
        self.check(self.lib.mipm_batch_dot(self.h, _ptr(x), _ptr(y), _ptr(out)))
        return out

    def batch_init_point_stage(self, stage, a=None, b=None, kappa=0.0):
        out = np.zeros((self._nb, 5))
        a = None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        b = None if b is None else np.ascontiguousarray(b, dtype=np.float64)
        self.check(self.lib.mipm_batch_init_point_stage(self.h, C.c_int(stage), _ptr(a), _ptr(b), C.c_double(kappa), _ptr(out)))
        return out

    def batch_iter_begin(self, del_w, del_c):
        out = np.zeros((self._nb, 16))
        st = C.c_int()
        self.check(self.lib.mipm_batch_iter_begin(self.h, C.c_double(del_w), C.c_double(del_c), _ptr(out), C.byref(st)))
        return out, st.value == MIPM_OK

    def batch_peek(self):
        out = np.zeros((self._nb, 16))
        self.check(self.lib.mipm_batch_peek(self.h, _ptr(out)))
        return out

    def batch_iter_rest(self, mu_min, step_rule, tau_param, ir_steps):
        self.check(self.lib.mipm_batch_iter_rest(self.h, C.c_double(mu_min), C.c_int(step_rule), C.c_double(tau_param), C.c_int(ir_steps)))

    def mpc_bind(self, vec: MpcVectors):
        self.check(self.lib.mipm_mpc_bind(self.h, C.byref(vec)))

    def set_aug_diagonal_reg(self, del_w, del_c):
        self.check(self.lib.mipm_set_aug_diagonal_reg(self.h, C.c_double(del_w), C.c_double(del_c)))

    def set_aug_diagonal_reg_scaled(self, del_w, del_c, sf):
        self.check(self.lib.mipm_set_aug_diagonal_reg_scaled(self.h, C.c_double(del_w), C.c_double(del_c), _ptr(sf)))

    def k25_scale_values(self, hi, hj, hraw, hout, jj, jraw, jout, sf, index_base=0):
        self.check(self.lib.mipm_k25_scale_values(self.h, C.c_int64(hraw.numel()), _ptr(hi), _ptr(hj), _ptr(hraw), _ptr(hout),
                                                  C.c_int64(jraw.numel()), _ptr(jj), _ptr(jraw), _ptr(jout), C.c_int(index_base), _ptr(sf)))

    def reduce_rhs_scaled(self, w, sf):
        self.check(self.lib.mipm_reduce_rhs_scaled(self.h, _ptr(w), _ptr(sf)))

    def finish_aug_solve_scaled(self, w, sf):
        self.check(self.lib.mipm_finish_aug_solve_scaled(self.h, _ptr(w), _ptr(sf)))

    def kktmul_scaled(self, w, v, alpha, beta):
        self.check(self.lib.mipm_kktmul_scaled(self.h, _ptr(w), _ptr(v), C.c_double(alpha), C.c_double(beta)))

    def set_predictive_rhs(self):
        self.check(self.lib.mipm_set_predictive_rhs(self.h))

    def set_correction_rhs(self, mu):
        self.check(self.lib.mipm_set_correction_rhs(self.h, C.c_double(mu)))

    def get_correction(self):
        self.check(self.lib.mipm_get_correction(self.h))

    def set_extra_correction(self, ap, ad, bmin, bmax, mu):
        self.check(self.lib.mipm_set_extra_correction(self.h, C.c_double(ap), C.c_double(ad), C.c_double(bmin),
                                                      C.c_double(bmax), C.c_double(mu)))

    def get_complementarity_measure(self):
        out = C.c_double()
        self.check(self.lib.mipm_get_complementarity_measure(self.h, C.byref(out)))
        return out.value

    def get_affine_complementarity_measure(self, ap, ad):
        out = C.c_double()
        self.check(self.lib.mipm_get_affine_complementarity_measure(self.h, C.c_double(ap), C.c_double(ad), C.byref(out)))
        return out.value

    def get_alpha_max(self, tau):
        a = (C.c_double * 4)()
        i = (C.c_int64 * 4)()
        self.check(self.lib.mipm_get_alpha_max(self.h, C.c_double(tau), a, i))
        return list(a), list(i)

    def mehrotra_adaptive_step(self, gamma_f):
        out = (C.c_double * 2)()
        self.check(self.lib.mipm_mehrotra_adaptive_step(self.h, C.c_double(gamma_f), out))
        return out[0], out[1]

    def termination_measures(self):
        out = (C.c_double * 5)()
        self.check(self.lib.mipm_termination_measures(self.h, out))
        return list(out)

    def apply_step(self, ap, ad, mu):
        self.check(self.lib.mipm_apply_step(self.h, C.c_double(ap), C.c_double(ad), C.c_double(mu)))

    def reduce_rhs(self, w):
        self.check(self.lib.mipm_reduce_rhs(self.h, _ptr(w)))

    def finish_aug_solve(self, w):
        self.check(self.lib.mipm_finish_aug_solve(self.h, _ptr(w)))

    def normal_solve_stage(self, stage, w, buffer_n, buffer_m):
        self.check(self.lib.mipm_normal_solve_stage(self.h, C.c_int(stage), _ptr(w), _ptr(buffer_n), _ptr(buffer_m)))

    def kktmul(self, w, v, alpha, beta):
        self.check(self.lib.mipm_kktmul(self.h, _ptr(w), _ptr(v), C.c_double(alpha), C.c_double(beta)))

    def residual_norms(self, w, p):
        out = (C.c_double * 2)()
        self.check(self.lib.mipm_residual_norms(self.h, _ptr(w), _ptr(p), out))
        return out[0], out[1]

    def init_point_stage(self, stage, a=0.0, b=0.0, kappa=0.0):
        out = (C.c_double * 8)()
        self.check(self.lib.mipm_init_point_stage(self.h, C.c_int(stage), C.c_double(a), C.c_double(b), C.c_double(kappa), out))
        return list(out)

    def init_bounds(self, n, tol, bound_push, bound_fac, x, xl, xu):
        self.check(self.lib.mipm_init_bounds(self.h, C.c_int64(n), C.c_double(tol), C.c_double(bound_push), C.c_double(bound_fac),
                                             _ptr(x), _ptr(xl), _ptr(xu)))

    def amax(self, n, x):
        out = C.c_double(0.0)
        self.check(self.lib.mipm_amax(self.h, C.c_int64(n), _ptr(x), C.byref(out)))
        return out.value

    def axpby(self, n, alpha, x, beta, y):
        self.check(self.lib.mipm_axpby(self.h, C.c_int64(n), C.c_double(alpha), _ptr(x), C.c_double(beta), _ptr(y)))

    def fill(self, n, value, x):
        self.check(self.lib.mipm_fill(self.h, C.c_int64(n), C.c_double(value), _ptr(x)))

    def copy(self, n, src, dst):
        self.check(self.lib.mipm_copy(self.h, C.c_int64(n), _ptr(src), _ptr(dst)))

    def gather(self, n, src, map_, dst, index_base=0):
        self.check(self.lib.mipm_gather(self.h, C.c_int64(n), _ptr(src), _ptr(map_), C.c_int(index_base), _ptr(dst)))

    def scatter(self, n, src, map_, dst, index_base=0):
        self.check(self.lib.mipm_scatter(self.h, C.c_int64(n), _ptr(src), _ptr(map_), C.c_int(index_base), _ptr(dst)))

    def dot(self, n, x, y):
        out = C.c_double()
        self.check(self.lib.mipm_dot(self.h, C.c_int64(n), _ptr(x), _ptr(y), C.byref(out)))
        return out.value

    def ls_factorize_profile(self, nzval):
        """Per-class device timing of one factorization: dict class -> (ms, work, launches)."""
        self._nz_ref = nzval
        ms, work, cnt = (C.c_double * 8)(), (C.c_double * 8)(), (C.c_int64 * 8)()
        self.check(self.lib.mipm_ls_factorize_profile(self.h, _ptr(nzval), ms, work, cnt))
        # classes 1-5: CTA-busy milliseconds summed over the grid / grid size; "kernel" = span of the task kernel
        names = ["zero_scatter", "extend_add", "diag", "trsm", "update", "wait", "kernel", "grid"]
        return {nm: dict(ms=ms[i], work=work[i], launches=cnt[i]) for i, nm in enumerate(names)}

    def launch_count(self):
        return int(self.lib.mipm_launch_count(self.h))

    def bench_syrk(self, n, k, Cmat, ldc, X, ldx):
        self.check(self.lib.mipm_bench_syrk(self.h, C.c_int64(n), C.c_int64(k), _ptr(Cmat), C.c_int64(ldc), _ptr(X), C.c_int64(ldx)))


def coo_to_csr(n_rows, n_cols, Ai, Aj, index_base=0):
    """mipm_coo_to_csr: returns (Bp, Bj, Bmap)."""
    Ai = np.ascontiguousarray(Ai, dtype=np.int32)
    Aj = np.ascontiguousarray(Aj, dtype=np.int32)
    nnz = len(Ai)
    Bp = np.zeros(n_rows + 1, dtype=np.int32)
    Bj = np.zeros(max(nnz, 1), dtype=np.int32)
    Bm = np.zeros(max(nnz, 1), dtype=np.int64)
    rc = load().mipm_coo_to_csr(C.c_int64(n_rows), C.c_int64(n_cols), C.c_int64(nnz), _ptr(Ai), _ptr(Aj),
                                C.c_int(index_base), _ptr(Bp), _ptr(Bj), _ptr(Bm))
    if rc != MIPM_OK:
        raise MipmError(rc, "mipm_coo_to_csr")
    return Bp, Bj[:nnz], Bm[:nnz]
