/*
 * oracle/sparse_ref.c -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * Plain-C restatement of the sparse-matrix functions on MadIPM's hot path, written from
 * the reference's Julia source (klamike/MadIPM.jl, read-only at /root/reference):
 *
 *   ref_coo_to_csr              <- src/utils.jl:158-201   (stable counting sort by row)
 *   ref_build_normal_system     <- src/utils.jl:209-274   (symbolic tril(A A^T), O(m^2) row scans)
 *   ref_assemble_normal_system  <- src/utils.jl:276-308   (numeric A D A^T over the fixed pattern)
 *   ref_transfer                <- MadNLP.transfer! (un-vendored MadNLP 0.8.12; semantics from
 *                                  ext/MadIPMCUDAExt/cuda_wrapper.jl:4-24: zero, then dest[map[k]] += src[k])
 *   ref_ldl_*                   <- the up-looking sparse LDL^T of T. Davis ("Algorithm 849: a concise
 *                                  sparse Cholesky factorization package", ACM TOMS 2005), which is the
 *                                  published algorithm LDLFactorizations.jl 0.10.1 (Project.toml:28,
 *                                  selected by test/runtests.jl:128,185 as `LDLSolver`) implements.
 *                                  That dependency is not vendored under /root/reference, so the
 *                                  algorithm is restated from the paper, not from source.
 *
 * All indices here are 0-based; the Python wrapper (oracle/sparse_ref.py) converts from the
 * reference's 1-based Int32 convention. The loops deliberately keep the reference's order of
 * floating-point operations (no FMA contraction: build with -ffp-contract=off) so the values are
 * the ones the reference's CPU path would produce.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* src/utils.jl:158-201. Bp has n_rows+1 entries; Bj/Bx have nnz entries. */
void ref_coo_to_csr(int64_t n_rows, int64_t nnz, const int32_t *Ai, const int32_t *Aj,
                    const double *Ax, int32_t *Bp, int32_t *Bj, double *Bx)
{
    for (int64_t i = 0; i <= n_rows; ++i) Bp[i] = 0;
    for (int64_t n = 0; n < nnz; ++n) Bp[Ai[n]] += 1;
    int32_t cumsum = 0;
    for (int64_t i = 0; i < n_rows; ++i) { int32_t t = Bp[i]; Bp[i] = cumsum; cumsum += t; }
    Bp[n_rows] = (int32_t)nnz;
    for (int64_t n = 0; n < nnz; ++n) {
        int32_t i = Ai[n];
        int32_t dest = Bp[i];
        Bj[dest] = Aj[n];
        Bx[dest] = Ax[n];
        Bp[i] += 1;
    }
    int32_t last = 0;
    for (int64_t i = 0; i <= n_rows; ++i) { int32_t t = Bp[i]; Bp[i] = last; last = t; }
}

/* src/utils.jl:209-274, pass 1 (count) when Cj == NULL, pass 2 (fill) otherwise.
 * Returns nnz. Cp gets per-row counts in pass 1 (caller prefix-sums), untouched in pass 2. */
int64_t ref_build_normal_system(int64_t n_rows, int64_t n_cols, const int32_t *Jtp,
                                const int32_t *Jtj, int32_t *Cp, int32_t *Cj)
{
    uint8_t *xb = (uint8_t *)calloc((size_t)(n_cols > 0 ? n_cols : 1), 1);
    int64_t nnz = 0;
    for (int64_t i = 0; i < n_rows; ++i) {
        for (int32_t c = Jtp[i]; c < Jtp[i + 1]; ++c) xb[Jtj[c]] = 1;
        for (int64_t j = i; j < n_rows; ++j) {
            for (int32_t c = Jtp[j]; c < Jtp[j + 1]; ++c) {
                if (xb[Jtj[c]] == 1) {
                    if (Cj) Cj[nnz] = (int32_t)j; else Cp[i] += 1;
                    nnz += 1;
                    break;
                }
            }
        }
        for (int32_t c = Jtp[i]; c < Jtp[i + 1]; ++c) xb[Jtj[c]] = 0;
    }
    free(xb);
    return nnz;
}

/* src/utils.jl:276-308. */
void ref_assemble_normal_system(int64_t n_rows, int64_t n_cols, const int32_t *Jtp,
                                const int32_t *Jtj, const double *Jtx, const int32_t *Cp,
                                const int32_t *Cj, double *Cx, const double *Dx)
{
    double *buffer = (double *)calloc((size_t)(n_cols > 0 ? n_cols : 1), sizeof(double));
    for (int64_t i = 0; i < n_rows; ++i) {
        for (int32_t c = Jtp[i]; c < Jtp[i + 1]; ++c) {
            int32_t j = Jtj[c];
            buffer[j] = Jtx[c] * Dx[j];
        }
        for (int32_t c = Cp[i]; c < Cp[i + 1]; ++c) {
            int32_t j = Cj[c];
            double acc = 0.0;
            for (int32_t d = Jtp[j]; d < Jtp[j + 1]; ++d) {
                double t = buffer[Jtj[d]] * Jtx[d];
                acc = acc + t;
            }
            Cx[c] = acc;
        }
        for (int32_t c = Jtp[i]; c < Jtp[i + 1]; ++c) buffer[Jtj[c]] = 0.0;
    }
    free(buffer);
}

/* MadNLP.transfer!(dest, src, map): fill!(dest,0); dest[map[k]] += src[k] in COO order. */
void ref_transfer(int64_t nnz_dest, double *dest, int64_t nnz_src, const double *src,
                  const int64_t *map)
{
    for (int64_t k = 0; k < nnz_dest; ++k) dest[k] = 0.0;
    for (int64_t k = 0; k < nnz_src; ++k) dest[map[k]] += src[k];
}

/* ---- Up-looking LDL^T (Davis 2005). Input: UPPER triangular CSC (Ap, Ai, Ax) of the
 * symmetric matrix, optional fill-reducing permutation P with inverse Pinv (NULL = identity).
 * ref_ldl_symbolic: elimination tree Parent and column counts Lnz, column pointers Lp. */
void ref_ldl_symbolic(int64_t n, const int64_t *Ap, const int32_t *Ai, int64_t *Lp,
                      int32_t *Parent, int64_t *Lnz, int32_t *Flag, const int32_t *P,
                      const int32_t *Pinv)
{
    for (int64_t k = 0; k < n; ++k) {
        Parent[k] = -1;
        Flag[k] = (int32_t)k;
        Lnz[k] = 0;
        int64_t kk = P ? P[k] : k;
        for (int64_t p = Ap[kk]; p < Ap[kk + 1]; ++p) {
            int32_t i = Pinv ? Pinv[Ai[p]] : Ai[p];
            if (i < k) {
                for (; Flag[i] != k; i = Parent[i]) {
                    if (Parent[i] == -1) Parent[i] = (int32_t)k;
                    Lnz[i]++;
                    Flag[i] = (int32_t)k;
                }
            }
        }
    }
    Lp[0] = 0;
    for (int64_t k = 0; k < n; ++k) Lp[k + 1] = Lp[k] + Lnz[k];
}

/* ref_ldl_numeric: returns n on success, else the index k of the first zero pivot D[k]==0.
 * Y (double n), Pattern/Flag (int32 n), Lnz (int64 n) are workspaces. */
int64_t ref_ldl_numeric(int64_t n, const int64_t *Ap, const int32_t *Ai, const double *Ax,
                        const int64_t *Lp, const int32_t *Parent, int64_t *Lnz, int32_t *Li,
                        double *Lx, double *D, double *Y, int32_t *Pattern, int32_t *Flag,
                        const int32_t *P, const int32_t *Pinv)
{
    for (int64_t k = 0; k < n; ++k) {
        Y[k] = 0.0;
        int64_t top = n;
        Flag[k] = (int32_t)k;
        Lnz[k] = 0;
        int64_t kk = P ? P[k] : k;
        for (int64_t p = Ap[kk]; p < Ap[kk + 1]; ++p) {
            int32_t i = Pinv ? Pinv[Ai[p]] : Ai[p];
            if (i <= k) {
                Y[i] += Ax[p];
                int64_t len;
                for (len = 0; Flag[i] != k; i = Parent[i]) {
                    Pattern[len++] = i;
                    Flag[i] = (int32_t)k;
                }
                while (len > 0) Pattern[--top] = Pattern[--len];
            }
        }
        D[k] = Y[k];
        Y[k] = 0.0;
        for (; top < n; ++top) {
            int32_t i = Pattern[top];
            double yi = Y[i];
            Y[i] = 0.0;
            int64_t p2 = Lp[i] + Lnz[i];
            int64_t p;
            for (p = Lp[i]; p < p2; ++p) Y[Li[p]] -= Lx[p] * yi;
            double l_ki = yi / D[i];
            D[k] -= l_ki * yi;
            Li[p] = (int32_t)k;
            Lx[p] = l_ki;
            Lnz[i]++;
        }
        if (D[k] == 0.0) return k;
    }
    return n;
}

/* Solve (P^T L D L^T P) x = b in place: X = b on entry, x on exit. W is a workspace of n. */
void ref_ldl_solve(int64_t n, double *X, const int64_t *Lp, const int32_t *Li,
                   const double *Lx, const double *D, const int32_t *P, double *W)
{
    double *B = X;
    if (P) { for (int64_t j = 0; j < n; ++j) W[j] = B[P[j]]; } else { memcpy(W, B, (size_t)n * sizeof(double)); }
    for (int64_t j = 0; j < n; ++j)
        for (int64_t p = Lp[j]; p < Lp[j + 1]; ++p) W[Li[p]] -= Lx[p] * W[j];
    for (int64_t j = 0; j < n; ++j) W[j] /= D[j];
    for (int64_t j = n - 1; j >= 0; --j)
        for (int64_t p = Lp[j]; p < Lp[j + 1]; ++p) W[j] -= Lx[p] * W[Li[p]];
    if (P) { for (int64_t j = 0; j < n; ++j) X[P[j]] = W[j]; } else { memcpy(X, W, (size_t)n * sizeof(double)); }
}
