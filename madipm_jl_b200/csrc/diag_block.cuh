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
constexpr int DLD = 68;         // shared-memory leading dimension (68 % 16 == 4: conflict-free DMMA fragment loads)
constexpr int PW = 16;          // panel width

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

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

// Rank-16 update of the trailing block on the FP64 tensor pipe: C(r, c) -= sum_j Y(r, j) / d_j * Y(c, j) over the lower
// 8 x 8 tiles of the nbelow x nbelow block (21 / 10 / 3 tiles for 48 / 32 / 16 rows). `next_diag`: only the three tiles of
// the next 16 x 16 diagonal block (they are on the critical path: the next panel factorization reads them); otherwise all
// the other tiles. Tiles are dealt round-robin to the `nw` worker warps (`wi` = index of this warp among them).
// Fragment convention of mma.m8n8k4: a = A[g][l3], b = B[l3][g], c[e] = C[g][2 l3 + e] with g = lane / 4, l3 = lane % 4.
__device__ __forceinline__ void trailing_tiles(double *S, const double *Ysc, const double *invd, int pc, int nbelow, int lane,
                                               int wi, int nw, bool next_diag)
{
    const int g = lane >> 2, l3 = lane & 3;
    const int nt = nbelow >> 3, g0 = pc + PW;
    int turn = 0;                                               // (no integer division: this sits on the critical path)
    for (int ti = next_diag ? 0 : 2; ti < (next_diag ? min(nt, 2) : nt); ++ti)
        for (int tj = 0; tj <= ti; ++tj) {
            const bool mine = (turn == wi);
            if (++turn == nw) turn = 0;
            if (!mine) continue;
            const double *ya = Ysc + pc + 8 * ti + g;           // Y(g0 + 8 ti + g, .)
            const double *yb = Ysc + pc + 8 * tj + g;           // Y(g0 + 8 tj + g, .)
            double a[PW / 4], b[PW / 4];
#pragma unroll
            for (int q = 0; q < PW / 4; ++q) {
                const int j = 4 * q + l3;
                a[q] = ya[j * DLD] * invd[pc + j];
                b[q] = yb[j * DLD];
            }
            double c0 = 0.0, c1 = 0.0;
#pragma unroll
            for (int q = 0; q < PW / 4; ++q) dmma884(c0, c1, a[q], b[q]);
            const int row = g0 + 8 * ti + g, col = g0 + 8 * tj + 2 * l3;
            if (row >= col) S[col * DLD + row] -= c0;
            if (row >= col + 1) S[(col + 1) * DLD + row] -= c1;
        }
}

// Block row ib >= 1 of inv(L_unit) (the diagonal blocks are in place):  M_ij = -M_ii sum_{k=16j}^{16i-1} L_ik M_kj, both
// products on the FP64 tensor pipe. 8 x 8 output tiles (2 x 2 ib of them) dealt to the nw worker warps; T = L M goes
// through the destination rows of Sinv between the two products. `sync` is the barrier of the worker set.
template <typename SyncF>
__device__ __forceinline__ void inverse_block_row(const double *S, double *Sinv, int ib, int lane, int wi, int nw, SyncF sync)
{
    const int g = lane >> 2, l3 = lane & 3;
    const int row0 = ib * PW, ntile = 4 * ib;
    double t0[2], t1[2];
    // T(row0 + r, c) = sum_{k = 16 (c / 16)}^{row0 - 1} L(row0 + r, k) M(k, c): tile = (ti, tc), widest K first
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int tile = wi + nw * u;
        t0[u] = t1[u] = 0.0;
        if (tile < ntile) {
            const int ti = tile & 1, tc = tile >> 1;
            const double *la = S + row0 + 8 * ti + g;                    // L(row0 + 8 ti + g, k) at la[k * DLD]
            const double *mb = Sinv + (8 * tc + g) * DLD;               // M(k, 8 tc + g) at mb[k]
            for (int k = PW * (tc >> 1); k < row0; k += 4) dmma884(t0[u], t1[u], la[(k + l3) * DLD], mb[k + l3]);
        }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int tile = wi + nw * u;
        if (tile < ntile) {
            const int ti = tile & 1, tc = tile >> 1;
            Sinv[(8 * tc + 2 * l3) * DLD + row0 + 8 * ti + g] = t0[u];
            Sinv[(8 * tc + 2 * l3 + 1) * DLD + row0 + 8 * ti + g] = t1[u];
        }
    }
    sync();
    // M(row0 + r, c) = -sum_{k' <= r} M_ii(r, k') T(row0 + k', c)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int tile = wi + nw * u;
        t0[u] = t1[u] = 0.0;
        if (tile < ntile) {
            const int ti = tile & 1, tc = tile >> 1;
            const double *ma = Sinv + row0 * DLD + row0 + 8 * ti + g;   // M_ii(8 ti + g, k') at ma[k' * DLD]
            const double *tb = Sinv + (8 * tc + g) * DLD + row0;        // T(row0 + k', 8 tc + g) at tb[k']
            for (int k = 0; k < 8 * (ti + 1); k += 4) dmma884(t0[u], t1[u], ma[(k + l3) * DLD], tb[k + l3]);
        }
    }
    sync();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int tile = wi + nw * u;
        if (tile < ntile) {
            const int ti = tile & 1, tc = tile >> 1;
            Sinv[(8 * tc + 2 * l3) * DLD + row0 + 8 * ti + g] = -t0[u];
            Sinv[(8 * tc + 2 * l3 + 1) * DLD + row0 + 8 * ti + g] = -t1[u];
        }
    }
}

// smem: [0, 64*68) S | [4352, 4352+256) dv, invd, spare | [4608, 4608+4352) Sinv (its upper right corner doubles as
// the panel scratch during the factorization).  Needs 8960 doubles.
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
    // Warp 0 owns the serial chain (the 16 x 16 panel factorizations). Everything else that panel q - 1 leaves behind --
    // the trailing update outside the next diagonal block and block row q - 1 of the inverse -- runs on warps 1-7 WHILE
    // warp 0 factors panel q, so the chain per panel is: factor -> rows below (all warps) -> next diagonal block.
    auto workers_sync = [] { asm volatile("bar.sync 1, 224;" ::: "memory"); };
#pragma unroll 1
    for (int pc = 0; pc < nbr; pc += PW) {
        if (warp != 0 && pc > 0) {
            trailing_tiles(S, Ysc, invd, pc - PW, nbr - pc, lane, warp - 1, 7, false);
            if (pc >= 2 * PW) inverse_block_row(S, Sinv, pc / PW - 1, lane, warp - 1, 7, workers_sync);
        }
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
#ifdef DIAG_EXP_TIME
        long long tt0 = clock64();
#endif
        if (warp < 3) trailing_tiles(S, Ysc, invd, pc, nbelow, lane, warp, 3, true);      // the next diagonal block only
#ifdef DIAG_EXP_TIME
        if (DBG && dbg) tq[4] += clock64() - tt0;
#endif
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

    // ---- the last block row of inv(L_unit) (the earlier ones were computed in the shadow of the panel factorizations)
    if (nbr > PW) {
        inverse_block_row(S, Sinv, nbr / PW - 1, lane, warp, 8, [] { __syncthreads(); });
        __syncthreads();
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
