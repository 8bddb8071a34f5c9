/*
 * madipm_b200.h -- C ABI of the B200-native (sm_100a) replacement for MadIPM.jl's
 * per-iteration Mehrotra predictor-corrector linear-algebra hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / CUDA.jl
 * types. A Julia `ext/MadIPMB200Ext` module binds these with `ccall`;
 * the Python host mirror binds them with ctypes. Each entry point cites the reference
 * interface (file:line under klamike/MadIPM.jl) it replaces.
 *
 * Conventions
 *   - `d_` pointers are DEVICE pointers (FP64 or index arrays as noted); others are host.
 *   - every entry that takes index arrays carries `index_base` (1 from Julia, 0 from C/Python).
 *   - all device work is enqueued on the stream given to mipm_create; the library only
 *     synchronises when it returns a host scalar (status, norms, step lengths).
 *   - return value: MIPM_OK or an error code; mipm_last_error(h) gives the message.
 *     Nothing throws, nothing aborts. There is no CPU fallback: without a CUDA device
 *     mipm_create fails with MIPM_ERR_CUDA.
 *   - host arrays returned through `T **out` are owned by the library until mipm_free(ptr).
 */
#ifndef MADIPM_B200_H
#define MADIPM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mipm_handle_s *mipm_handle;

enum {
    MIPM_OK = 0,
    MIPM_ERR_ARG = 1,            /* bad argument (null pointer, negative size, index out of range) */
    MIPM_ERR_CUDA = 2,           /* CUDA runtime error or no device                                */
    MIPM_ERR_ALLOC = 3,          /* host or device allocation failed                               */
    MIPM_ERR_STATE = 4,          /* call order violated (e.g. solve before factorize)              */
    MIPM_ERR_DUPLICATE = 5,      /* duplicate (row,col) entry where the reference assumes none     */
    MIPM_ERR_NOT_FACTORIZED = 6  /* pivot breakdown: maps to is_factorized(ls) == false            */
};

/* cudss_algorithm analogue (test/test_gpu.jl:9-11, scripts/benchmarks_gpu.jl:42). MIPM_LDL is for the quasi-definite K2
 * system (its analysis delays the dual vertices). MIPM_LDL_DEFINITE: the square-root-free LDL^T kernels on a matrix that is
 * positive definite up to rounding (the normal equations with cudss_algorithm = LDL): ordering as for Cholesky, and a
 * non-positive pivot late in an ill-conditioned solve is not a breakdown. */
enum { MIPM_CHOLESKY = 0, MIPM_LDL = 1, MIPM_LDL_DEFINITE = 2 };
enum { MIPM_ORDER_ND = 0, MIPM_ORDER_NATURAL = 1, MIPM_ORDER_USER = 2 };

/* ------------------------------------------------------------------ lifecycle ------ */
int mipm_version(void);
/* device: CUDA ordinal; stream: a cudaStream_t (NULL = legacy default stream). */
int mipm_create(mipm_handle *out, int device, void *stream);
int mipm_destroy(mipm_handle h);
const char *mipm_last_error(mipm_handle h);
void mipm_free(void *host_ptr);

/* ------------------------------------------------------------------ host symbolic -- */
/* Replaces MadIPM.coo_to_csr, src/utils.jl:158-207 (GPU override cuda_wrapper.jl:96-106):
 * stable counting sort of COO triplets by row. Bp[n_rows+1], Bj[nnz] use index_base;
 * Bmap[nnz] is the source position of each CSR entry in index_base too, i.e. what
 * NormalKKTSystem gets by pushing V = 1:nnz through coo_to_csr (normalkkt.jl:84-88). Host only. */
int mipm_coo_to_csr(int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t *Ai,
                    const int32_t *Aj, int index_base, int32_t *Bp, int32_t *Bj, int64_t *Bmap);

/* Replaces MadIPM.build_normal_system, src/utils.jl:209-274 (GPU: cuda_wrapper.jl:158-234):
 * pattern of tril(A A') as lower CSC (column i holds rows j >= i, ascending). (Ap, Aj) is
 * the CSR of A (= colptr/rowval of AT, normalkkt.jl:92). Bit-exact with the reference's
 * O(m^2) scan but O(sum of row merges). Also builds, inside the handle, the product-term
 * map that mipm_normal_assemble consumes. *Cp (m+1) and *Cj (nnzC) are library-owned host
 * arrays in index_base (release with mipm_free). Rows of A must not repeat a column
 * (MIPM_ERR_DUPLICATE). A handle with a GPU builds both on the device (one CTA per row of the
 * pattern sorts the row's product terms in shared memory); Ap == Aj == NULL then means "the
 * matrix registered with mipm_spmv_setup on this handle" (same m, n), whose CSR / CSC index
 * is already on the device. Analysis-only handles (device < 0) run the host sweep. */
int mipm_normal_symbolic(mipm_handle h, int64_t m, int64_t n, const int32_t *Ap,
                         const int32_t *Aj, int index_base, int32_t **Cp, int32_t **Cj,
                         int64_t *nnzC);

/* Replaces MadNLP.compress_jacobian!(::NormalKKTSystem), normalkkt.jl:163-172 /
 * cuda_wrapper.jl:32-41, for the assembly side: d_ATx = AT.nzVal (CSR order of A, nnz(A)
 * doubles). J is constant over the solve (solver.jl:167), so the per-term weights
 * A[i,k]*A[j,k] are precomputed here once. */
int mipm_normal_set_jacobian(mipm_handle h, const double *d_ATx);

/* Replaces build_kkt!(::NormalKKTSystem) + assemble_normal_system!, normalkkt.jl:180-194,
 * utils.jl:276-308, cuda_wrapper.jl:108-156: d_Cx[c] = sum_k A[i,k] * (1/pr_diag[k]) * A[j,k]
 * over the fixed pattern. exact_order != 0 reproduces the reference CPU loop's operation
 * order bit for bit ((A[i,k]*D[k])*A[j,k] summed in row-j order, no FMA). */
int mipm_normal_assemble(mipm_handle h, const double *d_pr_diag, double *d_Cx, int exact_order);

/* Replaces MadNLP.SparseKKTSystem's coo_to_csc at construction (un-vendored MadNLP 0.8.12;
 * call site src/utils.jl:110): lower-triangular COO (I >= J) of dimension dim -> lower CSC
 * (rows ascending, duplicates merged, SparseArrays.sparse pattern) + map coo -> csc slot. */
int mipm_k2_symbolic(mipm_handle h, int64_t dim, int64_t nnz_coo, const int32_t *I,
                     const int32_t *J, int index_base, int32_t **colptr, int32_t **rowval,
                     int64_t **map, int64_t *nnz_csc);

/* Replaces MadNLP.transfer!(aug_com, aug_raw, aug_csc_map) = K2 build_kkt!
 * (cuda_wrapper.jl:4-24): d_nz = 0; d_nz[map[k]] += d_V[k]. Done as a deterministic gather
 * per CSC slot in COO order (the reference kernel races on duplicates). */
int mipm_k2_transfer(mipm_handle h, const double *d_V, double *d_nz);

/* ------------------------------------------------------------------ linear solver -- */
/* Replaces the MadNLP.AbstractLinearSolver constructor `linear_solver(aug_com; opt)`
 * (normalkkt.jl:113-115; cuDSS `analysis` inside MadNLPGPU.CUDSSSolver): host ordering,
 * elimination tree, supernodes, assembly maps, level schedule; device workspace allocation.
 * Matrix: lower-triangular CSC, n x n, pattern fixed for the handle's lifetime. */
int mipm_ls_analyze(mipm_handle h, int64_t n, const int32_t *colptr, const int32_t *rowval,
                    int index_base, int kind, int ordering, const int32_t *user_perm);

/* Replaces MadNLP.factorize!(linear_solver) (cuDSS factorization/refactorization) and
 * MadIPM.is_factorized (src/utils.jl:54-62): numeric supernodal multifrontal Cholesky /
 * LDL^T of the values d_nzval (same order as rowval). *status = MIPM_OK or
 * MIPM_ERR_NOT_FACTORIZED (non-positive / tiny pivot); the call itself returns MIPM_OK then. */
int mipm_ls_factorize(mipm_handle h, const double *d_nzval, int *status);
/* Same, but leaves the status on the device; mipm_ls_status reads it back later (one sync). */
int mipm_ls_factorize_async(mipm_handle h, const double *d_nzval);
int mipm_ls_status(mipm_handle h, int *status);

/* Replaces MadNLP.solve!(linear_solver, x) (normalkkt.jl:210): in-place solve of K x = b with
 * `ir_steps` rounds of iterative refinement against the d_nzval given to the last factorize. */
int mipm_ls_solve(mipm_handle h, double *d_x, int ir_steps);

/* MadNLP.inertia(linear_solver) analogue (signs of the pivots of the last factorization). */
int mipm_ls_inertia(mipm_handle h, int64_t *num_pos, int64_t *num_zero, int64_t *num_neg);

typedef struct {
    int64_t n;
    int64_t nnz_a;           /* nnz of the lower triangle handed in                       */
    int64_t nnz_l;           /* nonzeros of L incl. diagonal (supernodal storage, padded) */
    int64_t nnz_l_exact;     /* sum of column counts                                      */
    double flops;            /* sum_j colcount_j^2                                        */
    int64_t n_supernodes;
    int64_t n_levels;
    int64_t max_front_cols;  /* widest supernode                                          */
    int64_t max_front_rows;  /* tallest front (cols + below rows)                        */
    int64_t update_doubles;  /* total size of the update (Schur) matrices                 */
    int64_t n_launches;      /* kernel launches per numeric factorization                 */
} mipm_ls_stats_t;
int mipm_ls_stats(mipm_handle h, mipm_ls_stats_t *out);
/* Caps the grid of the persistent factorization / solve kernels of this handle (0 = whole GPU, the default). Call before
 * mipm_ls_analyze. With a small cap several handles on different streams run side by side on one GPU: the batch of small
 * independent LPs of BASELINE config C5 (the reference solves such batches one after the other). */
int mipm_set_grid_limit(mipm_handle h, int max_ctas);

/* Symbolic structure export (for parity tests: "GPU library == oracle" on the same
 * deterministic host analysis). All arrays 0-based, library-owned.
 *   perm[n]: new -> old;  sn_ptr[ns+1]: first column of each supernode;
 *   sn_parent[ns];  row_ptr[ns+1] / row_idx: below-diagonal row structure per supernode. */
int mipm_ls_symbolic(mipm_handle h, int32_t **perm, int64_t *n_sn, int32_t **sn_ptr,
                     int32_t **sn_parent, int64_t **row_ptr, int32_t **row_idx);

/* ---- distributed (block-angular) solves: SURVEY 8e, config C4. New functionality (the reference has no
 * multi-GPU code). Each rank analyses and factors ITS OWN sub-matrix [interior blocks of its commodities; border]
 * whose last n_border rows/columns (the linking constraints) are kept last as one dense root supernode:
 *   mipm_ls_analyze_border           host analysis with that constraint
 *   mipm_ls_factorize_stage(.., 0)   assembly + every front below the root: the root panel then holds this rank's
 *                                    Schur contribution (the last extend-add wrote straight into it)
 *   [caller: all-reduce (sum) of the n_root x n_root root panel over NCCL]
 *   mipm_ls_factorize_stage(.., 1)   factorization of the (now global) root front, redundantly on every rank
 *   mipm_ls_solve_stage(x, 0)        forward sweep below the root; the root segment of the permuted RHS is partial
 *   [caller: all-reduce (sum) of the n_root root RHS entries]
 *   mipm_ls_solve_stage(x, 1)        root solve + backward sweep; x then holds this rank's part of the solution
 *   mipm_ls_root_info                device pointers to the root panel (column-major, ld = n_root) and root RHS. */
int mipm_ls_analyze_border(mipm_handle h, int64_t n, const int32_t *colptr, const int32_t *rowval,
                           int index_base, int kind, int64_t n_border);
int mipm_ls_factorize_stage(mipm_handle h, const double *d_nzval, int stage);
int mipm_ls_solve_stage(mipm_handle h, double *d_x, int stage);
int mipm_ls_root_info(mipm_handle h, double **d_root_panel, int64_t *n_root, double **d_root_rhs);

/* ------------------------------------------------------------------ SpMV ----------- */
/* Replaces MadIPMOperator / cuSPARSE SpMV on AT (cuda_wrapper.jl:43-94; normalkkt.jl:177,
 * 208,214,228-229): CSR of A (m x n) registered once (host arrays), then
 *   trans == 0: y = alpha * A  x + beta * y   (y length m)
 *   trans == 1: y = alpha * A' x + beta * y   (y length n), via a CSC index built at setup. */
int mipm_spmv_setup(mipm_handle h, int64_t m, int64_t n, const int32_t *Ap, const int32_t *Aj,
                    int index_base);
int mipm_spmv(mipm_handle h, int trans, double alpha, const double *d_Ax, const double *d_x,
              double beta, double *d_y);
/* The two products of one phase in ONE launch: y1 = alpha1 A x1 + beta1 y1 (length m) and y2 = alpha2 A' x2 + beta2 y2
 * (length n). They are independent of each other in evaluate_model! (c = A x - b and jacl = A' y, src/solver.jl:319-326)
 * and in the residual of solve_system! / mul!(w, kkt, d) (w_y -= A d_x, w_x -= A' d_y: src/linear_solver.jl:29-35,
 * normalkkt.jl:221-233). Falls back to two mipm_spmv calls when no column-ordered copy of the values is cached. */
int mipm_spmv_pair(mipm_handle h, const double *d_Ax, double alpha1, const double *d_x1, double beta1, double *d_y1,
                   double alpha2, const double *d_x2, double beta2, double *d_y2);
/* Tells the library that the CSR values at d_Ax stay unchanged until the next call of this function (in MadIPM
 * compress_jacobian!, normalkkt.jl:163-172, is their only writer): a column-ordered copy is kept so that
 * mul!(y, AT', x) streams its values instead of gathering them through the position map. mipm_spmv uses the copy
 * only when it is called with the same d_Ax; d_Ax = NULL drops the copy. */
int mipm_spmv_cache_values(mipm_handle h, const double *d_Ax);

/* Hessian operator for obj / grad / the (1,1) block of mul!: replaces MadIPMOperator(H; symmetric=true),
 * cuda_wrapper.jl:62-68 (H expanded to its full symmetric CSR by the caller, as `tril(A,-1) + A'` does):
 * y = alpha * H x + beta * y with d_Hx the CSR-ordered values. */
int mipm_hess_setup(mipm_handle h, int64_t n, const int32_t *Hp, const int32_t *Hj, int index_base);
int mipm_hess_spmv(mipm_handle h, double alpha, const double *d_Hx, const double *d_x, double beta,
                   double *d_y);

/* ------------------------------------------------------------------ MPC vectors ---- */
/* Device buffers of one MPCSolver (src/structure.jl:1-77, 125-153) and of its KKT system
 * (normalkkt.jl:58-67). KKT vectors d, p, w are [xp(n); y(m); zl(nlb); zu(nub)]
 * (structure.jl:132). ind_lb/ind_ub are Int64 like Julia's Vector{Int}. */
typedef struct {
    int64_t n, m, nlb, nub;
    int index_base;
    const int64_t *d_ind_lb, *d_ind_ub;
    double *d_x, *d_xl, *d_xu, *d_zl, *d_zu, *d_f;     /* full(.) of the PrimalVectors, length n */
    double *d_y, *d_c, *d_rhs;                           /* length m */
    double *d_jacl;                                      /* length n */
    double *d_d, *d_p, *d_w;                             /* solver.d, solver.p, solver._w1 */
    double *d_corr_lb, *d_corr_ub;                       /* correction_lb / correction_ub */
    double *d_reg, *d_pr_diag, *d_du_diag;               /* kkt.reg, kkt.pr_diag, kkt.du_diag */
    double *d_l_diag, *d_u_diag, *d_l_lower, *d_u_lower; /* kkt bound blocks */
} mipm_mpc_vectors;
int mipm_mpc_bind(mipm_handle h, const mipm_mpc_vectors *v);

/* set_aug_diagonal_reg!, src/kernels.jl:124-136 (one fused launch instead of eight). */
int mipm_set_aug_diagonal_reg(mipm_handle h, double del_w, double del_c);
/* K2.5 = MadNLP.ScaledSparseKKTSystem (test/runtests.jl:107-120, test/test_gpu.jl:9): the K2 system scaled symmetrically by
 * S = diag(sqrt((x - xl)(xu - x))) (absent bounds count as 1) so that its diagonal stays bounded as bounds become active.
 *   mipm_set_aug_diagonal_reg_scaled  set_aug_diagonal_reg!(::ScaledSparseKKTSystem), src/kernels.jl:139-149 +
 *                                     MadNLP._set_aug_diagonal!: l_diag = x - xl, u_diag = xu - x (sign flipped against K2),
 *                                     pr_diag = zu (x - xl) + zl (xu - x) + reg S^2, d_scaling_factor = S (length n)
 *   mipm_k25_scale_values             build_kkt!: hess_out = hess_raw * S_i S_j, jac_out = jac_raw * S_col (device index arrays)
 *   mipm_reduce_rhs_scaled / mipm_finish_aug_solve_scaled   solve!: reduce_rhs! with the positive diagonals, primal block
 *                                     times S before and after the linear solve, then the bound duals
 *   mipm_kktmul_scaled                _kktmul! for the residual check on the unscaled unreduced system */
int mipm_set_aug_diagonal_reg_scaled(mipm_handle h, double del_w, double del_c, double *d_scaling_factor);
int mipm_k25_scale_values(mipm_handle h, int64_t nnzh, const int32_t *d_hess_i, const int32_t *d_hess_j,
                          const double *d_hess_raw, double *d_hess_out, int64_t nnzj, const int32_t *d_jac_j,
                          const double *d_jac_raw, double *d_jac_out, int index_base, const double *d_scaling_factor);
int mipm_reduce_rhs_scaled(mipm_handle h, double *d_w, const double *d_scaling_factor);
int mipm_finish_aug_solve_scaled(mipm_handle h, double *d_w, const double *d_scaling_factor);
int mipm_kktmul_scaled(mipm_handle h, double *d_w, const double *d_v, double alpha, double beta);
/* set_predictive_rhs! / set_correction_rhs!, src/kernels.jl:21-58. */
int mipm_set_predictive_rhs(mipm_handle h);
int mipm_set_correction_rhs(mipm_handle h, double mu);
/* get_correction! / set_extra_correction!, src/kernels.jl:60-122. */
int mipm_get_correction(mipm_handle h);
int mipm_set_extra_correction(mipm_handle h, double alpha_p, double alpha_d, double beta_min,
                              double beta_max, double mu);
/* get_complementarity_measure / get_affine_complementarity_measure, src/kernels.jl:155-208. */
int mipm_get_complementarity_measure(mipm_handle h, double *out);
int mipm_get_affine_complementarity_measure(mipm_handle h, double alpha_p, double alpha_d,
                                            double *out);
/* get_alpha_max_primal + get_alpha_max_dual, src/kernels.jl:226-272, in one launch.
 * alpha[4] = (alpha_xl, alpha_xu, alpha_zl, alpha_zu); idx[4] = 1-based arg-min positions
 * within the lb/ub blocks (0 = the init value 1.0 won), first minimum wins like the
 * reference's strict '<' reduction. get_fraction_to_boundary_step (kernels.jl:274-289) is
 * min(alpha[0],alpha[1]), min(alpha[2],alpha[3]). */
int mipm_get_alpha_max(mipm_handle h, double tau, double *alpha, int64_t *idx);
/* update_step!(::MehrotraAdaptiveStep), src/kernels.jl:309-358 (Mehrotra's GTSF heuristic), entirely on the device: the
 * four ratio tests with tau = 1, the affine complementarity measure at (alpha_p^max, alpha_d^max) and the element reads
 * at the blocking indices, which the reference does by scalar indexing into device arrays from the host.
 * alpha[2] = (alpha_p, alpha_d). One synchronisation (none through mipm_mpc_iter_rest with step_rule = 2). */
int mipm_mehrotra_adaptive_step(mipm_handle h, double gamma_f, double *alpha);
/* dual_objective (kernels.jl:408-417), get_inf_pr / get_inf_du / get_optimality_gap
 * (solver.jl:196-204, kernels.jl:419-430) and ||primal(d)||_inf (structure.jl:193) in one
 * launch. out[5] = (dobj, ||c||_inf, ||f - zl + zu + jacl||_inf, max compl, ||dx||_inf). */
int mipm_termination_measures(mipm_handle h, double *out);
/* apply_step! incl. MadNLP.adjust_boundary!, src/solver.jl:308-317. */
int mipm_apply_step(mipm_handle h, double alpha_p, double alpha_d, double mu);
/* MadNLP.reduce_rhs! / finish_aug_solve! on a KKT vector (normalkkt.jl:197,217). */
int mipm_reduce_rhs(mipm_handle h, double *d_w);
int mipm_finish_aug_solve(mipm_handle h, double *d_w);
/* The vector part of solve!(::NormalKKTSystem, w), normalkkt.jl:196-219, around the two SpMVs
 * and the linear solve (buffer_n = kkt.buffer_n, buffer_m = kkt.buffer_m):
 *   stage 0 (:197-207): reduce_rhs!(w); buffer_n = wx ./ Sigma; buffer_m = wy
 *       [host: buffer_m = A*buffer_n - buffer_m (mipm_spmv); mipm_ls_solve(buffer_m)]
 *   stage 1 (:212-213): wy = buffer_m; buffer_n = wx
 *       [host: buffer_n -= A'*wy (mipm_spmv)]
 *   stage 2 (:215-217): wx = buffer_n ./ Sigma; finish_aug_solve!(w). */
int mipm_normal_solve_stage(mipm_handle h, int stage, double *d_w, double *d_buffer_n,
                            double *d_buffer_m);
/* MadNLP._kktmul! (the bound/regularization part of mul!, normalkkt.jl:231):
 * w <- w (already alpha*K_sparse*v + beta*w in its primal/dual part) + alpha*reg.*vx etc. */
int mipm_kktmul(mipm_handle h, double *d_w, const double *d_v, double alpha, double beta);
/* norm(full(w), Inf), norm(full(p), Inf) of solve_system!, src/linear_solver.jl:32-33.
 * out[2] = (||w||_inf, ||p||_inf); NaN propagates like Julia's norm. */
int mipm_residual_norms(mipm_handle h, const double *d_w, const double *d_p, double *out);
/* init_starting_point! vector statements, src/solver.jl:41-118 (after the two solves):
 * stage 0: zl/zu init from res = jacl + f (res passed in d_jacl), returns mins for delta_x/delta_s
 * in out[4] = (min(x_lr-xl_r), min(xu_r-x_ur), min zl_r, min zu_r);
 * stage 1: shifts by (delta_x, delta_s) then returns out[5] = (mu, sum zl_r, sum zu_r,
 * sum(x_lr-xl_r), sum(xu_r-x_ur)); stage 2: shifts by (delta_x2, delta_s2) + projection with kappa,
 * returns out[4] = interior check mins (min zl_r, min zu_r, min(x_lr-xl_r), min(xu_r-x_ur)). */
int mipm_init_point_stage(mipm_handle h, int stage, double a, double b, double kappa, double *out);
/* MadNLP.initialize! bound handling (called from MadIPM.initialize!, src/solver.jl:127-140), in place
 * on device vectors of length n that hold the raw values on entry:
 *   xl -= max(1,|xl|)*tol;  xu += max(1,|xu|)*tol                                  (bound_relax_factor)
 *   x   = projection of x strictly inside [xl, xu] with bound_push / bound_fac     (initialize_variables!)
 * Infinite bounds stay infinite. Products and sums are rounded separately (no FMA) so the result is
 * bit-identical to the reference's broadcast statements. */
int mipm_init_bounds(mipm_handle h, int64_t n, double tol, double bound_push, double bound_fac,
                     double *d_x, double *d_xl, double *d_xu);
/* out = max_i |x_i| (0 for n = 0): norm(x, Inf) as used for norm_b / norm_c and the gradient scaling. */
int mipm_amax(mipm_handle h, int64_t n, const double *d_x, double *out);
/* Plain fused helpers used by the host loop in place of broadcast statements:
 * y[i] = alpha*x[i] + beta*y[i];  fill; copy. */
int mipm_axpby(mipm_handle h, int64_t n, double alpha, const double *d_x, double beta, double *d_y);
int mipm_fill(mipm_handle h, int64_t n, double value, double *d_x);
int mipm_copy(mipm_handle h, int64_t n, const double *d_src, double *d_dst);
/* dst[i] = src[map[i] - index_base], i < n: the device gather behind
 * `kkt.AT.nzVal .= kkt.A.V[kkt.A_csr_map]` (cuda_wrapper.jl:39). map is Int64 like Julia's Vector{Int}. */
int mipm_gather(mipm_handle h, int64_t n, const double *d_src, const int64_t *d_map, int index_base,
                double *d_dst);
/* dst[map[i] - index_base] = src[i], i < n (map entries must be distinct): used to place a rank's part of a
 * distributed solution into the global vector. */
int mipm_scatter(mipm_handle h, int64_t n, const double *d_src, const int64_t *d_map, int index_base,
                 double *d_dst);
/* dot(x, y) with a deterministic two-level reduction (used for obj = c'x + x'Qx/2). */
int mipm_dot(mipm_handle h, int64_t n, const double *d_x, const double *d_y, double *out);

/* ------------------------------------------------------------------ fused iteration - */
/* One mpc! iteration (src/solver.jl:332-360) with device-resident scalars: the step lengths, the
 * centering parameter and the barrier value never leave the GPU, so an iteration costs ONE host
 * synchronisation (termination measures + factorization status) instead of ~20.
 * Same kernels and operation order as the fine-grained entry points above.
 *   mipm_mpc_set_model : which KKT system (0 = NormalKKTSystem, 1 = K2 SparseKKTSystem) and its buffers.
 *   mipm_mpc_iter_begin: update_termination_criteria! measures + set_aug_diagonal_reg! + build_kkt! +
 *                        factorize!, then one sync. out[16] = (dobj, ||c||inf, ||f-zl+zu+jacl||inf,
 *                        max compl, ||dx||inf, obj c'x, x'Hx, alpha_p, alpha_d, mu, mu_curr,
 *                        ||w||inf, ||p||inf of the predictor solve, same for the corrector solve, tau)
 *                        where entries 5.. describe the PREVIOUS iteration's step. *status as mipm_ls_factorize.
 *   mipm_mpc_refactor  : the retry of factorize_regularized_system! with new (del_w, del_c).
 *   mipm_mpc_iter_rest : prediction_step! + mehrotra_correction_direction! + update_step_size!
 *                        (step_rule 0 = AdaptiveStep(tau_param), 1 = ConservativeStep(tau_param),
 *                        2 = MehrotraAdaptiveStep(gamma_f = tau_param)) +
 *                        apply_step! + evaluate_model!; nothing is read back. */
typedef struct {
    int kkt_kind;                 /* 0 Normal, 1 K2 */
    int exact_order;              /* Normal: reference operation order in the assembly */
    int64_t nx;                   /* number of original variables (Hessian dimension) */
    double c0;                    /* objective constant (already scaled) */
    const double *d_ATx;          /* AT.nzVal (CSR values of A incl. slack columns) */
    const double *d_cvec;         /* scaled linear objective, length n */
    const double *d_Hx;           /* full symmetric CSR Hessian values or NULL */
    double *d_aug_nz;             /* aug_com.nzVal */
    const double *d_aug_raw_V;    /* K2: COO values [pr_diag; hess; jac; du_diag] */
    double *d_buffer_n, *d_buffer_m;
} mipm_mpc_model;
int mipm_mpc_set_model(mipm_handle h, const mipm_mpc_model *model);
int mipm_mpc_iter_begin(mipm_handle h, double del_w, double del_c, double *out, int *status);
/* The termination measures of the current iterate alone (update_termination_criteria!, src/solver.jl:194-205) into the
 * same 16 scalars as mipm_mpc_iter_begin, WITHOUT assembling and factorizing the next KKT system. The host calls it
 * instead of mipm_mpc_iter_begin when the previous measures are within 10 x tol: the iteration that detects convergence
 * then does not pay for a factorization it never uses (one extra synchronisation in the last one or two iterations). */
int mipm_mpc_peek(mipm_handle h, double *out);
int mipm_mpc_refactor(mipm_handle h, double del_w, double del_c, int *status);
int mipm_mpc_iter_rest(mipm_handle h, double mu_min, int step_rule, double tau_param, int ir_steps);

/* The same fused iteration around an EXTERNAL linear solver (the distributed block-angular solver, whose staged
 * factorization / solves have NCCL exchanges between them): the iteration is cut where the normal system is factorized
 * and solved. NormalKKTSystem only. Per iteration: _ext_begin (measures + diagonals + assembly) -> caller factorizes
 * aug_nz -> _ext_fetch (the one synchronisation, out[16] as mipm_mpc_iter_begin) -> _ext_phase(0) -> caller solves
 * buffer_m in place -> _ext_phase(1) -> caller solves buffer_m -> _ext_phase(2). */
int mipm_mpc_ext_begin(mipm_handle h, double del_w, double del_c);
int mipm_mpc_ext_fetch(mipm_handle h, double *out);
int mipm_mpc_ext_phase(mipm_handle h, int phase, double mu_min, int step_rule, double tau_param);

/* ------------------------------------------------------------------ batches ---------- */
/* BASELINE config C5 (a batch of independent LPs / QPs; the reference solves them one after the other). The B units
 * are STACKED into one block-diagonal problem -- x, bounds, multipliers and KKT vectors of unit u occupy
 * [off_n[u], off_n[u+1]) of the variable blocks and [off_m[u], off_m[u+1]) of the row blocks, A and H are block
 * diagonal -- and bound / analysed exactly like a single problem (mipm_mpc_bind, mipm_spmv_setup,
 * mipm_normal_symbolic, mipm_ls_analyze: the elimination forest has one tree per unit). Element-wise kernels, SpMVs,
 * assembly, factorization and solves then run ONCE per IPM phase for all units; the entry points below do what is
 * per unit (reductions, step lengths, centering, barrier value, termination) with one CTA per unit:
 *   mipm_batch_configure           unit offsets (host arrays of B + 1), after mipm_mpc_bind
 *   mipm_batch_set_active          active[B] (host): a unit with active = 0 keeps its iterate (it has terminated)
 *   mipm_batch_amax / _dot         per-unit norms / dot products of stacked vectors (scaling, objective), out[B]
 *   mipm_batch_init_point_stage    mipm_init_point_stage per unit: a[B], b[B] in, out[B x 5]
 *   mipm_batch_iter_begin / _rest  mipm_mpc_iter_begin / mipm_mpc_iter_rest for the batch: out[B x 16]; the status
 *                                  is that of the stacked factorization (mipm_mpc_refactor retries for all units);
 *                                  no residual check (entries 11-14 are 0). */
int mipm_batch_configure(mipm_handle h, int64_t n_units, const int64_t *off_n, const int64_t *off_m);
int mipm_batch_set_active(mipm_handle h, const int32_t *active);
int mipm_batch_amax(mipm_handle h, int by_rows, const double *d_x, double *out);
int mipm_batch_dot(mipm_handle h, const double *d_x, const double *d_y, double *out);
int mipm_batch_init_point_stage(mipm_handle h, int stage, const double *a, const double *b, double kappa,
                                double *out);
int mipm_batch_iter_begin(mipm_handle h, double del_w, double del_c, double *out, int *status);
/* mipm_mpc_peek for the whole batch: the termination measures of every unit without the next factorization. */
int mipm_batch_peek(mipm_handle h, double *out);
int mipm_batch_iter_rest(mipm_handle h, double mu_min, int step_rule, double tau_param, int ir_steps);

/* ------------------------------------------------------------------ preprocessing -- */
/* Ruiz equilibration on the device: replaces `Dr, Dc = HSL.mc77(A, 0)` of scale_qp (scripts/common.jl:57-100; TODO at
 * src/solver.jl:147). A is given in COO form (device arrays, index_base); on return A ./ (Dr_i Dc_j) has rows and
 * columns of infinity norm ~1. max_iter sweeps (MC77's default is 10), stopped early when tol > 0 and
 * max |1 - norm| <= tol. mipm_scale_coo applies the scaling to a COO value array (H with (Dc, Dc), A with (Dr, Dc)),
 * like _scale_coo! (scripts/common.jl:37-44). */
int mipm_ruiz_equilibrate(mipm_handle h, int64_t m, int64_t n, int64_t nnz, const int32_t *d_rows,
                          const int32_t *d_cols, const double *d_vals, int index_base, int max_iter, double tol,
                          double *d_Dr, double *d_Dc, int *iters);
int mipm_scale_coo(mipm_handle h, int64_t nnz, const int32_t *d_rows, const int32_t *d_cols,
                   const double *d_vals, int index_base, const double *d_Dr, const double *d_Dc, double *d_out);

/* ------------------------------------------------------------------ diagnostics ---- */
/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
int64_t mipm_launch_count(mipm_handle h);
/* One timed numeric factorization with the task kernel's own per-class accounting (all arrays have 8 entries):
 * ms[0] zero-fill + scatter of the input values (CUDA-event time minus the kernel span); ms[1..4] extend-add,
 * diagonal blocks (+ small leaf fronts), panel solves, trailing DMMA updates: %globaltimer busy time of the tasks of
 * that class summed over all CTAs and divided by the grid size; ms[5] the same for dependency waits; ms[6] the
 * span of the task kernel (first task start to last task end); ms[7] the grid size. work[] = algorithmic bytes
 * (classes 0-1) or flops (classes 2-4). MIPM_PHASE_LOG=<file> also writes when each elimination-tree level finished.
 * For bench.py's roofline. */
int mipm_ls_factorize_profile(mipm_handle h, const double *d_nzval, double *ms, double *work,
                              int64_t *launches);
/* Dense FP64 update-kernel micro-benchmark hook: C(n x n, lower tiles) -= X(n x k) X(n x k)'
 * with the same DMMA tile kernel the factorization uses; for roofline measurement. */
int mipm_bench_syrk(mipm_handle h, int64_t n, int64_t k, double *d_C, int64_t ldc,
                    const double *d_X, int64_t ldx);

#ifdef __cplusplus
}
#endif
#endif /* MADIPM_B200_H */
