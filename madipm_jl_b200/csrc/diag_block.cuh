// 64 x 64 diagonal block of a front: factorization + inverse of the triangular factor, one CTA of 256 threads.
// Replaces the dense diagonal-block work cuDSS does inside `factorize!` (reference call site:
// src/linear_solver.jl:10 via MadNLP.factorize_wrapper!). Shared by csrc/factor.cu (task_diag) and the
// stand-alone harness tools/diag_v2_bench.cu.
//
// Panel-blocked, square-root-free right-looking elimination (A = L D L', L unit lower):
//   for each 16-column panel
//     (1) ONE warp factors the 16 x 16 diagonal block in registers (row per lane, shuffles): the only serial
//         chain per column is shuffle -> reciprocal -> multiply -> FMA; its upper 16 lanes carry the columns of
//         the block's inverse through the same instruction stream, for free;
//     (2) the rows below are multiplied by that inverse (16-term dots, all threads);
//     (3) all threads apply the rank-16 update to the trailing block (4 x 2 register tiles, lower tiles only).
//   Then the off-diagonal block rows of inv(L) follow from two 16-wide products each. Cholesky scales by sqrt(D) on the way out (off the critical path).
// 12 CTA barriers for the factorization instead of one per column, and ~4x fewer issued instructions than the
// column-at-a-time version it replaces (which mattered twice: latency near the root of the tree, issue
// throughput in the wide bottom levels where three CTAs share an SM).
#pragma once
#include <cuda_runtime.h>

namespace mipm_diag {

constexpr int DB = 64;          // block size
constexpr int DLD = 65;         // shared-memory leading dimension
constexpr int PW = 16;          // panel width

// Branch-free reciprocal for a normal, finite, non-zero argument: MUFU seed (relative error 2^-23) and two Newton steps.
// (__drcp_rn ends in a slow-path branch that keeps ptxas from overlapping it with independent work.) Arguments outside
// that range are caught by pivot_bad and recomputed by the careful pass.
__device__ __forceinline__ double fast_rcp(double d)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    e = fma(-d, x, 1.0);
    x = fma(x, e, x);
    return x;
}

template <bool LDL>
__device__ __forceinline__ bool pivot_bad(double d, double piv_tol)
{
    if (!LDL) return !(d > 0.0) || !(d < 1.0e300);
    return !(fabs(d) <= 1.0e300) || fabs(d) < piv_tol;
}
// Cholesky: a non-positive pivot is flagged and replaced by 1 so the run stays finite (the host retries with more
// regularization, src/linear_solver.jl:6-17). LDL^T: |pivot| < piv_tol is replaced by +-piv_tol.
template <bool LDL>
__device__ __forceinline__ void pivot_fix(double &d, double piv_tol, int &nbad, int &ntiny)
{
    if (LDL && fabs(d) <= 1.0e300) { ntiny++; d = (d < 0.0) ? -piv_tol : piv_tol; }
    else { nbad++; d = 1.0; }
}

// Careful version of the warp-level 16 x 16 panel factorization (same layout as the fast pass in diag_block): pivots are
// tested and fixed one by one. Kept out of line so its registers do not weigh on the fast pass.
template <bool LDL>
__device__ __noinline__ void careful_panel(double *S, double *Sinv, double *dv, double *invd, int pc, double piv_tol,
                                           int &nbad, int &ntiny)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, i = lane & 15;
    const bool fac = lane < PW;
    double a[PW];
#pragma unroll
    for (int k = 0; k < PW; ++k) a[k] = fac ? S[(pc + k) * DLD + pc + i] : ((k == i) ? 1.0 : 0.0);
#pragma unroll
    for (int j = 0; j < PW; ++j) {
        double d = __shfl_sync(FULL, a[j], j);
        if (pivot_bad<LDL>(d, piv_tol)) pivot_fix<LDL>(d, piv_tol, nbad, ntiny);
        const double inv = 1.0 / d;
        const double aj = a[j];
        const double lij = aj * inv;
#pragma unroll
        for (int k = j + 1; k < PW; ++k) a[k] = fma(-lij, __shfl_sync(FULL, aj, k), a[k]);
        if (fac) {
            if (i > j) a[j] = lij;
            else if (i == j) a[j] = d;
        }
        if (lane == j) { dv[pc + j] = d; invd[pc + j] = inv; }
    }
    if (fac) {
#pragma unroll
        for (int k = 0; k < PW; ++k) if (k <= i) S[(pc + k) * DLD + pc + i] = a[k];
    } else {
#pragma unroll
        for (int k = 0; k < PW; ++k) Sinv[(pc + i) * DLD + pc + k] = a[k];
    }
}

// Rank-16 update of the trailing block (8 NCOL rows / columns). Lane = row (two halves of 32), warp w = columns
// w, w + 8, ...: the row operand is read conflict-free (consecutive rows), the column operand is a broadcast.
template <int NCOL>
__device__ __forceinline__ void trailing_update(double *S, const double *Ysc, const double *invd, int pc, int lane, int warp)
{
    constexpr int NBELOW = 8 * NCOL;
    constexpr bool TWO = NBELOW > 32;
    const bool h1 = TWO && (lane + 32) < NBELOW;
    const double *Yr = Ysc + pc + lane;
    double acc0[NCOL], acc1[NCOL];
#pragma unroll
    for (int t = 0; t < NCOL; ++t) acc0[t] = acc1[t] = 0.0;
    if (lane < NBELOW) {
#pragma unroll
        for (int j = 0; j < PW; ++j) {
            const double dj = invd[pc + j];
            const double y0 = Yr[j * DLD] * dj;
            const double y1 = h1 ? Yr[j * DLD + 32] * dj : 0.0;
#pragma unroll
            for (int t = 0; t < NCOL; ++t) {
                const double yc = Ysc[j * DLD + pc + warp + 8 * t];
                acc0[t] = fma(y0, yc, acc0[t]);
                if (TWO) acc1[t] = fma(y1, yc, acc1[t]);
            }
        }
        const int g0 = pc + PW;
#pragma unroll
        for (int t = 0; t < NCOL; ++t) {
            const int c = warp + 8 * t;
            if (lane >= c) S[(g0 + c) * DLD + g0 + lane] -= acc0[t];
            if (h1) S[(g0 + c) * DLD + g0 + lane + 32] -= acc1[t];
        }
    }
}

// smem: [0, 64*65) S | [4160, 4160+256) dv, invd, spare | [4416, 4416+4160) Sinv (its upper right corner doubles as
// the panel scratch during the factorization).  Needs 8576 doubles.
// P: the block inside the front panel (column-major, leading dimension N); nb <= 64 valid rows / columns.
// preloaded: S already holds the block (lower triangle inside nb, identity elsewhere); P is then only written.
// Dv: 64 x 64 column-major output with column stride ldv, inverse of the stored factor (identity outside nb).
// info[0] = breakdown flag, info[1] = # negative pivots (LDL^T), info[2] = # perturbed pivots.
template <bool LDL, bool DBG = false>
__device__ __forceinline__ void diag_block(double *P, int N, int nb, double *Dv, int ldv,
                                           double piv_tol, int *info, double *smem, bool preloaded = false, long long *dbg = nullptr)
{
    double *S = smem;
    double *dv = smem + DB * DLD;           // pivots d_j, later sqrt(d_j)
    double *invd = dv + DB;                 // 1/d_j, later 1/sqrt(d_j)
    double *cbuf = invd + DB;               // 2 x 16 column buffer of the panel factorization
    double *Sinv = smem + DB * DLD + 4 * DB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;

    long long tq[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tc = clock64();
#define DIAG_STAMP(i) do { if (DBG && dbg) { long long now_ = clock64(); tq[i] += now_ - tc; tc = now_; } } while (0)
    if (!preloaded) {                       // (preloaded: the caller already placed the block, identity-padded, in S)
        double v[DB * DB / 256];            // all 16 global loads in flight before the first shared store
#pragma unroll
        for (int t = 0; t < DB * DB / 256; ++t) {
            const int idx = tid + 256 * t, rr = idx & 63, cc = idx >> 6;
            v[t] = (rr < nb && cc < nb && rr >= cc) ? P[(int64_t)cc * N + rr] : ((rr == cc) ? 1.0 : 0.0);
        }
#pragma unroll
        for (int t = 0; t < DB * DB / 256; ++t) {
            const int idx = tid + 256 * t;
            S[(idx >> 6) * DLD + (idx & 63)] = v[t];
        }
    }
    __syncthreads();
    DIAG_STAMP(0);

    int nbad = 0, ntiny = 0;                // meaningful in warp 0 / lane 0 only
    double *Ysc = Sinv + (DB - PW) * DLD;   // panel scratch Y(r, j) at Ysc[j * DLD + r - PW]: the part of Sinv above its
                                            // last diagonal block, which the inverse never uses
    // Only the 16-column panels that hold real columns are factored: the rest of the block is the identity padding. Fronts
    // of K2 systems and of block-angular problems are mostly narrower than 32 columns (C4: 19 on average), so this is most
    // of the diagonal-block work there.
    const int nbr = min(DB, ((nb + PW - 1) / PW) * PW);
#pragma unroll 1
    for (int pc = 0; pc < nbr; pc += PW) {
        if (warp == 0) {
            // lanes 0-15: row i of the 16 x 16 block. lanes 16-31: column i of inv(L11), carried through the SAME
            // instruction stream: with m = e_i in place of the row, step j does m[k] -= (m[j]/d_j) A(k,j), which is
            // forward substitution with the unit-lower factor.
            const int i = lane & 15;
            const bool fac = lane < PW;
            double a[PW];
#pragma unroll
            for (int k = 0; k < PW; ++k) a[k] = fac ? S[(pc + k) * DLD + pc + i] : ((k == i) ? 1.0 : 0.0);
            // Serial chain per column: multiply -> FMA (lane j+1's own next pivot) -> shuffle -> reciprocal. The other
            // updates of column j are issued while the reciprocal of pivot j+1 is in flight. The fast pass has no
            // branch at all (ptxas does not move instructions across one, and a branch per column cost 40 cycles of
            // the chain: tools/panel16_bench.cu): pivot tests only accumulate a flag, and a panel that saw a bad
            // pivot is redone from shared memory by the careful pass (rare: it means the factorization broke down
            // or LDL^T needed a perturbation).
            bool redo = false;
            {
                double d = __shfl_sync(FULL, a[0], 0);
                double inv = fast_rcp(d);
                redo = pivot_bad<LDL>(d, piv_tol);
#pragma unroll
                for (int j = 0; j < PW; ++j) {
                    const double aj = a[j];                          // A(i, j) before scaling
                    const double lij = aj * inv;
                    // column j goes through shared memory (one store, broadcast loads): a 64-bit shuffle is two
                    // instructions per operand and this warp is issue-bound
                    double *cj = cbuf + (j & 1) * PW;
                    if (fac) cj[i] = aj;
                    double d_n = 1.0, inv_n = 1.0;
                    if (j + 1 < PW) {
                        const double own = fma(-lij, aj, a[j + 1]);  // exact for lane j+1: its updated diagonal entry
                        d_n = __shfl_sync(FULL, own, j + 1);
                        inv_n = fast_rcp(d_n);
                        redo = redo || pivot_bad<LDL>(d_n, piv_tol);
                    }
                    __syncwarp();
#pragma unroll
                    for (int k = j + 1; k < PW; ++k) a[k] = fma(-lij, cj[k], a[k]);
                    if (fac) {
                        if (i > j) a[j] = lij;
                        else if (i == j) a[j] = d;
                    }
                    if (lane == j) { dv[pc + j] = d; invd[pc + j] = inv; }
                    d = d_n; inv = inv_n;
                }
            }
            if (redo) {                                              // warp-uniform, rare
                __syncwarp();
                careful_panel<LDL>(S, Sinv, dv, invd, pc, piv_tol, nbad, ntiny);
            } else if (fac) {
#pragma unroll
                for (int k = 0; k < PW; ++k) if (k <= i) S[(pc + k) * DLD + pc + i] = a[k];
            } else {
#pragma unroll
                for (int k = 0; k < PW; ++k) Sinv[(pc + i) * DLD + pc + k] = a[k];      // M(k, i), zero above the diagonal
            }
        }
        const int nbelow = nbr - pc - PW;
        if (nbelow == 0) break;             // uniform
        __syncthreads();
        DIAG_STAMP(1);
        // rows below: Y = A21 inv(L11)' (16-term dots), L21 = Y D^-1. Thread (r, jq): row r, columns 4 jq .. 4 jq + 3
        {
            const int r = tid & 63, jq = tid >> 6;
            double l4[4] = {0.0, 0.0, 0.0, 0.0};
            const bool below = r >= pc + PW && r < nbr;
            if (below) {
                double x[PW];
#pragma unroll
                for (int k = 0; k < PW; ++k) x[k] = S[(pc + k) * DLD + r];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = 4 * jq + jj;
                    double y0 = 0.0, y1 = 0.0;
#pragma unroll
                    for (int k = 0; k < PW; k += 2) {
                        y0 = fma(x[k], Sinv[(pc + k) * DLD + pc + j], y0);
                        y1 = fma(x[k + 1], Sinv[(pc + k + 1) * DLD + pc + j], y1);
                    }
                    const double y = y0 + y1;
                    Ysc[j * DLD + r - PW] = y;
                    l4[jj] = y * invd[pc + j];
                }
            }
            __syncthreads();
            DIAG_STAMP(2);
            if (below) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) S[(pc + 4 * jq + jj) * DLD + r] = l4[jj];
            }
        }
        // trailing block -= Y D^-1 Y'
        if (nbelow == 48) trailing_update<6>(S, Ysc, invd, pc, lane, warp);
        else if (nbelow == 32) trailing_update<4>(S, Ysc, invd, pc, lane, warp);
        else trailing_update<2>(S, Ysc, invd, pc, lane, warp);
        __syncthreads();
        DIAG_STAMP(3);
    }
    __syncthreads();
    DIAG_STAMP(1);
    if (tid == 0) {
        if (nbad) atomicMax(&info[0], 1);
        if (ntiny) atomicAdd(&info[2], ntiny);
    }
    if (LDL && tid < nb && dv[tid] < 0.0) atomicAdd(&info[1], 1);

    // ---- inv(L_unit), block rows i = 1..3 (the diagonal blocks are in place):  M_ij = -M_ii sum_{k=j}^{i-1} L_ik M_kj.
    // Thread (r, c0) owns row 16 i + r of the columns c0 + 16 jb, jb < i.
    {
        const int r = tid & 15, c0 = tid >> 4;
#pragma unroll
        for (int ib = 1; ib < DB / PW; ++ib) {
            if (ib * PW >= nbr) break;                                   // uniform: identity padding needs no inverse
            const int row0 = ib * PW;
            double t[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int jb = 0; jb < 3; ++jb) {
                if (jb < ib) {
                    const int c = c0 + PW * jb;
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                    for (int k = PW * jb; k < row0; k += 4) {           // block-aligned start: M is zero above its diagonal
                        s0 = fma(S[k * DLD + row0 + r], Sinv[c * DLD + k], s0);
                        s1 = fma(S[(k + 1) * DLD + row0 + r], Sinv[c * DLD + k + 1], s1);
                        s2 = fma(S[(k + 2) * DLD + row0 + r], Sinv[c * DLD + k + 2], s2);
                        s3 = fma(S[(k + 3) * DLD + row0 + r], Sinv[c * DLD + k + 3], s3);
                    }
                    t[jb] = (s0 + s1) + (s2 + s3);
                }
            }
#pragma unroll
            for (int jb = 0; jb < 3; ++jb) if (jb < ib) Sinv[(c0 + PW * jb) * DLD + row0 + r] = t[jb];
            __syncthreads();
#pragma unroll
            for (int jb = 0; jb < 3; ++jb) {
                if (jb < ib) {
                    const int c = c0 + PW * jb;
                    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                    for (int k = 0; k < PW; k += 4) {                   // M_ii is zero above its diagonal
                        s0 = fma(Sinv[(row0 + k) * DLD + row0 + r], Sinv[c * DLD + row0 + k], s0);
                        s1 = fma(Sinv[(row0 + k + 1) * DLD + row0 + r], Sinv[c * DLD + row0 + k + 1], s1);
                        s2 = fma(Sinv[(row0 + k + 2) * DLD + row0 + r], Sinv[c * DLD + row0 + k + 2], s2);
                        s3 = fma(Sinv[(row0 + k + 3) * DLD + row0 + r], Sinv[c * DLD + row0 + k + 3], s3);
                    }
                    t[jb] = -((s0 + s1) + (s2 + s3));
                }
            }
            __syncthreads();
#pragma unroll
            for (int jb = 0; jb < 3; ++jb) if (jb < ib) Sinv[(c0 + PW * jb) * DLD + row0 + r] = t[jb];
            __syncthreads();
        }
    }
    DIAG_STAMP(5);
    // ---- write out: the factor (Cholesky: L = L_unit sqrt(D); LDL^T: unit multipliers, D on the diagonal) and its
    // inverse (Cholesky: D^-1/2 inv(L_unit))
    if (tid < DB) {
        const double d = (tid < nbr) ? dv[tid] : 1.0;
        const double sq = LDL ? d : sqrt(d);
        dv[tid] = sq;
        invd[tid] = LDL ? 1.0 : 1.0 / sq;
    }
    __syncthreads();
    for (int idx = tid; idx < DB * DB; idx += 256) {
        const int rr = idx & 63, cc = idx >> 6;
        if (rr >= cc && rr < nb && cc < nb) {
            const double v = S[cc * DLD + rr];
            P[(int64_t)cc * N + rr] = (rr == cc) ? dv[cc] : (LDL ? v : v * dv[cc]);
        }
        if (rr >= nbr || cc >= nbr) Dv[cc * ldv + rr] = (rr == cc) ? 1.0 : 0.0;       // identity outside the factored panels
        else Dv[cc * ldv + rr] = (rr >= cc) ? Sinv[cc * DLD + rr] * invd[rr] : 0.0;
    }
    DIAG_STAMP(6);
    if (DBG && dbg && tid == 0) for (int i = 0; i < 7; ++i) dbg[i] = tq[i];
#undef DIAG_STAMP
}

}  // namespace mipm_diag
